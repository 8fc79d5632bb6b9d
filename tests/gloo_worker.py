"""torchrun worker (CPU, gloo): every rank plans ITS shard with the native packer/scheduler
(plan-only handle, no GPU), executes the plan with the numpy interpreter and performs the one
exchange step as a real torch.distributed all_gather.  Checks the sharded result against the
oracle.  Launched by tests/test_sharded_gloo.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import hssb200 as hb  # noqa: E402
import hss_oracle as o  # noqa: E402
import plan_interp as pi  # noqa: E402


def sharded_product(rank, world, n, ls, r, seed, k, adjoint):
    """One sharded product executed by the interpreter; adjoint=True runs the forward plan over the adjoint
    twin pool (A' X on a sharded handle, hssb_matmul_t)."""
    P = hb.synthetic(n, ls, r, seed, shard_rank=rank, n_shards=world, plan_only=True)
    X = o.synth_x(seed, n, k, P.info.local_col0, P.info.local_n)
    Y = np.full((P.info.local_m, k), np.nan, order="F")
    st = pi.ShardState(P, X, Y, k, P.debug_pool_t() if adjoint else None)
    for ph in [p for p in st.phases if not p.transposed]:
        if ph.kind == pi.PH_EXCHANGE:  # all-gather of the subtree-root Z blocks, in place in the Z workspace
            cnt, base = ph.xchg_slot_rows * k, ph.xchg_zoff * k
            mine = torch.from_numpy(st.Z[base + rank * cnt: base + (rank + 1) * cnt].copy())
            slots = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(slots, mine)
            for g, sl in enumerate(slots):
                st.Z[base + g * cnt: base + (g + 1) * cnt] = sl.numpy()
        else:
            st.run_phase(ph, 1.0, 0.0)
    h = o.synthetic_hss(n, ls, r, seed)
    ref = o.matmul(o.adjoint(h) if adjoint else h, o.synth_x(seed, n, k))
    mine = ref[P.info.local_row0:P.info.local_row0 + P.info.local_m]
    return np.linalg.norm(Y - mine) / np.linalg.norm(mine)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    e = max(sharded_product(rank, world, 1024, 64, 6, 9, 5, False),      # any-shape layout
            sharded_product(rank, world, 2048, 128, 16, 11, 3, False),   # padded (fixed-shape kernel) layout
            sharded_product(rank, world, 2048, 128, 16, 11, 3, True))    # ... and its adjoint twin pool
    err = torch.tensor([e])
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    dist.destroy_process_group()
    if rank == 0:
        print("GLOO_SHARDED_OK" if err.item() <= 1e-12 else f"GLOO_SHARDED_FAIL {err.item():.3e}")
    sys.exit(0 if err.item() <= 1e-12 else 1)


if __name__ == "__main__":
    main()
