"""ncu tensor-pipe metrics per GEMM class -> profiles/<tag>_gemm_tensor.json.

    python profiles/gemm_shapes.py --ncu > gpurun_out/gemm_order.txt          (plain run first)
    ncu --metrics <METRICS> --clock-control none -k regex:gemm --csv --log-file gpurun_out/gemm_tensor.csv \
        python profiles/gemm_shapes.py --ncu
    python profiles/summarize_gemm_tensor.py gpurun_out/gemm_tensor.csv gpurun_out/gemm_order.txt r2a

Every shape is launched twice (warm-up + measured); the second launch of each pair is kept.
"""
import collections
import csv
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_tensor.sum", "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32.sum",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(csv_path, order_path, tag):
    order = [l.split(" ", 1)[1].strip().split("|") for l in open(order_path) if l.startswith("NCU_ORDER")]
    rows = list(csv.reader(open(csv_path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi, mi, ii, ui = (h.index(x) for x in ("Kernel Name", "Metric Value", "Metric Name", "ID", "Metric Unit"))
    d = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        e = d.setdefault(r[ii], {"kernel": r[ki].split("(")[0]})
        try:
            e[r[mi]] = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
        except ValueError:
            pass
    launches = [e for e in d.values() if "gemm" in e["kernel"]]
    assert len(launches) == 2 * len(order), (len(launches), len(order))
    out = []
    for i, (name, M, K, N, flops) in enumerate(order):
        e = launches[2 * i + 1]
        us = e["gpu__time_duration.sum"]
        rec = {"gemm": name, "M": int(M), "K": int(K), "N": int(N), "kernel": e["kernel"].replace("void hmocr::<unnamed>::", ""),
               "time_us_under_ncu": round(us, 1), "tflops_algorithmic": round(float(flops) / us / 1e6, 1)}
        for k in METRICS[1:]:
            if k in e:
                rec[k] = e[k]
        ops = e.get("sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32.sum")
        if ops:
            rec["tensor_path_tflops"] = round(ops / us / 1e6, 1)      # ops counted by the tensor path itself / duration
        out.append(rec)
    json.dump({"how": "ncu --metrics ... --clock-control none, one launch per GEMM class of the Swin-T encoder at B=256 "
                      "(profiles/gemm_shapes.py --ncu); times under ncu are cold-cache and serialised",
               "gemms": out}, open(os.path.join(HERE, f"{tag}_gemm_tensor.json"), "w"), indent=1)
    for r in out:
        print(r["gemm"], r["time_us_under_ncu"], r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
              r["tflops_algorithmic"])


if __name__ == "__main__":
    main(*sys.argv[1:4])
