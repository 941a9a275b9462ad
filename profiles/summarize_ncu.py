"""Turn ncu outputs in gpurun_out/ into the small, committed evidence files of profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches_r1g.csv r1g   # launch list of the last step -> profiles/r1g_launches.csv,
                                                                                # decode DRAM traffic -> profiles/decode_traffic.json
    python profiles/summarize_ncu.py full gpurun_out/dp_r1g.ncu-rep r1g_decode_persistent   # --set full capture -> metrics json + stall summary
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void hmocr::<unnamed>::", "").replace("hmocr::<unnamed>::", "")


def kernel_source_sha():
    """Hash of the decode kernel's sources at capture time: bench.py prints `roofline.traffic` from decode_traffic.json
    only while the library it runs was built from the same sources."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(os.path.dirname(HERE), "handwritten_math_ocr_api_b200", "csrc")
    for f in ("decode_persistent.cu", "decode_persistent.cuh"):
        h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()


def launches(path, tag, batch=256, max_len=150):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi, mi, ii, ui = (h.index(x) for x in ("Kernel Name", "Metric Value", "Metric Name", "ID", "Metric Unit"))
    d = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        e = d.setdefault(r[ii], {"name": short(r[ki])})
        e[r[mi]] = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
    items = list(d.values())
    start = max(i for i, e in enumerate(items) if "patch_embed" in e["name"])      # last generate() call
    step = items[start:]
    out = os.path.join(HERE, f"{tag}_launches.csv")
    tot = sum(e["gpu__time_duration.sum"] for e in step)
    with open(out, "w") as f:
        f.write("# one generate() call, B=%d, T=%d: ncu --metrics gpu__time_duration.sum,dram__bytes_* --clock-control none\n"
                % (batch, max_len))
        f.write("idx,kernel,time_us,share_of_step,dram_read_MB,dram_write_MB\n")
        for i, e in enumerate(step):
            f.write("%d,%s,%.1f,%.4f,%.1f,%.1f\n" % (i, e["name"], e["gpu__time_duration.sum"],
                                                     e["gpu__time_duration.sum"] / tot,
                                                     e.get("dram__bytes_read.sum", 0) / 1e6,
                                                     e.get("dram__bytes_write.sum", 0) / 1e6))
    dec = [e for e in step if "decode_persistent_kernel" in e["name"]]
    traffic = sum(e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0) for e in dec)
    t_dec = sum(e["gpu__time_duration.sum"] for e in dec)
    agg = collections.OrderedDict()
    for e in step:
        a = agg.setdefault(e["name"], [0, 0.0])
        a[0] += 1
        a[1] += e["gpu__time_duration.sum"]
    summary = {"batch": batch, "max_len": max_len, "source": os.path.basename(path),
               "kernel_source_sha256": kernel_source_sha(),
               "decode_kernel_launches": len(dec), "decode_kernel_time_us_under_ncu": t_dec,
               "dram_bytes_per_step": traffic, "step_time_us_under_ncu": tot,
               "decode_kernel_share_of_step": t_dec / tot,
               "by_kernel": {k: {"launches": v[0], "time_us": round(v[1], 1), "share": round(v[1] / tot, 4)}
                             for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
    json.dump(summary, open(os.path.join(HERE, "decode_traffic.json"), "w"), indent=1)
    print(json.dumps(summary, indent=1))


WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_uniform.sum"]


def full(rep, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, vals = rows[0], rows[1], rows[2]
    m = {"kernel": vals[h.index("Kernel Name")], "source": os.path.basename(rep)}
    for i, name in enumerate(h):
        if name in WANT:
            m[name] = {"value": vals[i], "unit": units[i]}
    stalls = {}
    for i, name in enumerate(h):
        if name.startswith("smsp__average_warps_issue_stalled") and name.endswith("per_issue_active.ratio") and vals[i]:
            stalls[name.split("stalled_")[1].split("_per")[0]] = float(vals[i].replace(",", ""))
    m["warps_stalled_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
    json.dump(m, open(os.path.join(HERE, f"{tag}_ncu_metrics.json"), "w"), indent=1)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h = rows[1]
    c = {x: i for i, x in enumerate(h)}
    k = "Warp Stall Sampling (All Samples)"
    data, tot = [], 0
    for idx, r in enumerate(rows[2:]):
        try:
            v = int(r[c[k]].replace(",", ""))
        except Exception:
            continue
        tot += v
        data.append((v, idx, r))
    reasons = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    rt = {x: 0 for x in reasons}
    for v, idx, r in data:
        for x in reasons:
            try:
                rt[x] += int(r[c[x]].replace(",", ""))
            except Exception:
                pass
    with open(os.path.join(HERE, f"{tag}_stalls.txt"), "w") as f:
        f.write(f"# warp-stall sampling of {m['kernel'][:80]} ({os.path.basename(rep)}), {tot} samples, {len(data)} SASS instructions\n")
        f.write("# by reason:\n")
        for x, v in sorted(rt.items(), key=lambda kv: -kv[1])[:10]:
            f.write(f"  {x:22s} {100.0 * v / max(tot, 1):5.1f} %\n")
        f.write("# top 40 instructions (samples, share, SASS index, instruction):\n")
        for v, idx, r in sorted(data, key=lambda x: -x[0])[:40]:
            f.write(f"  {v:7d} {100.0 * v / tot:5.1f}% #{idx:5d} {r[c['Source']][:90]}\n")
    print(json.dumps(m, indent=1)[:3000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3])
