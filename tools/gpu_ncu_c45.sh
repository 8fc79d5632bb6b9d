#!/bin/bash
# ncu summaries of one product on the shapes of configs 4 and 5 (one GPU's share at 4 resp. 8 GPUs: 8192 leaves each): per launch
# device time, DRAM bytes, DRAM %, FP64 tensor path %
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread
for cfg in "c4 1048576 128 64 128" "c5 2097152 256 64 32"; do
  set -- $cfg
  timeout 600 ncu --metrics $M --clock-control none -k regex:"leaf2|node_kernel" -s 54 -c 27 --csv --log-file gpurun_out/ncu_$1.csv python tools/phases.py $2 $3 $4 $5 > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
done
python - <<'P'
import csv, collections, json
out = {}
for cfg, algo in (("c4", None), ("c5", None)):
    rows = [r for r in csv.reader(open(f"gpurun_out/ncu_{cfg}.csv")) if len(r) > 10]
    hdr = rows[0]
    d = collections.OrderedDict()
    for r in rows[1:]:
        rec = dict(zip(hdr, r))
        e = d.setdefault(rec["ID"], {"kernel": rec["Kernel Name"][:60], "grid": rec["Grid Size"]})
        e[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", ""))
    L = []
    for e in d.values():
        L.append({"kernel": e["kernel"], "grid": e["grid"], "us": e["gpu__time_duration.sum"] / 1e3,
                  "dram_bytes": e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"],
                  "dram_pct": e["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"],
                  "fp64_tensor_pct": e["sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed"], "regs": e["launch__registers_per_thread"]})
    out[cfg] = {"launches": L, "total_us_under_ncu": sum(x["us"] for x in L), "total_dram_bytes": sum(x["dram_bytes"] for x in L)}
    print(cfg, "total us", round(out[cfg]["total_us_under_ncu"], 1), "dram GB", round(out[cfg]["total_dram_bytes"] / 1e9, 3))
    for x in L:
        if x["us"] > 20: print("   %-50s %9.1f us  %7.3f GB  dram %5.1f %%  fp64 tensor %5.1f %%" % (x["kernel"], x["us"], x["dram_bytes"] / 1e9, x["dram_pct"], x["fp64_tensor_pct"]))
json.dump(out, open("gpurun_out/ncu_c4_c5_kernels.json", "w"), indent=1)
P
