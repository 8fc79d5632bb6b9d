"""torchrun worker: one rank per GPU, subtree-sharded synthetic matrix, one
NCCL all-gather per product; compares the gathered result with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import hssb200 as hb  # noqa: E402
import hss_oracle as o  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    n, ls, r, k, seed = 8192, 128, 32, 64, 17
    hs = o.synthetic_hss(n, ls, r, seed)
    ref = o.matmul(hs, o.synth_x(seed, n, k))
    ref_t = o.matmul(o.adjoint(hs), o.synth_x(seed, n, k))
    err = 0.0
    for mode in ("nccl", "peer"):
        P = hb.synthetic(n, ls, r, seed, device=local, shard_rank=rank, n_shards=world)
        if mode == "nccl":      # one ncclAllGather per product
            uid = [hb.PackedHss.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            P.comm_init(uid[0], rank, world)
        else:                   # NVLink peer stores into IPC-mapped workspaces
            P.reserve(k)
            blobs = [None] * world
            dist.all_gather_object(blobs, P.xchg_export())
            P.xchg_import(blobs)
        rows = P.info.local_n
        X = o.synth_x(seed, n, k, P.info.local_col0, rows)
        mine = ref[P.info.local_row0:P.info.local_row0 + P.info.local_m]
        for rep in range(3):    # repeated products exercise the epoch / acknowledgement protocol
            Y = P @ X
            err = max(err, np.linalg.norm(Y - mine) / np.linalg.norm(mine))
        # A' X on the sharded handle: forward plan (same exchange) over the adjoint twin pool
        mine_t = ref_t[P.info.local_col0:P.info.local_col0 + P.info.local_n]
        for rep in range(2):
            Yt = P.tmatmul(X)
            err = max(err, np.linalg.norm(Yt - mine_t) / np.linalg.norm(mine_t))
        Y = P @ X
        err = max(err, np.linalg.norm(Y - mine) / np.linalg.norm(mine))
        dist.barrier()
        P.close()
    t = torch.tensor([err], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDED_OK" if t.item() <= 1e-12 else f"SHARDED_FAIL {t.item():.3e}")
    if t.item() > 1e-12:
        sys.exit(1)


if __name__ == "__main__":
    main()
