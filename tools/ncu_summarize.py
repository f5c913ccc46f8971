#!/usr/bin/env python
"""Turn the raw ncu output of tools/gpu_profile.sh (gpurun_out/) into the small tracked summaries under profiles/.

  python tools/ncu_summarize.py --round r01

* gpurun_out/launches_bench.csv (ncu --metrics gpu__time_duration.sum of the bench command)
    -> profiles/<round>_launches_bench_step.csv : per kernel launches / step, µs / step, share of the step
       (a "step" = one forward; the number of forwards in the capture = launches of sls_tail_kernel)
    -> profiles/<round>_launches_bench_raw.csv.gz
* gpurun_out/prof_{gemm,ln2,hbm,pool}.ncu-rep (ncu --set full) -> profiles/<round>_ncu_{gemm,conv_pair,hbm_kernels,sls_pool}_summary.csv
* profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch (bench.py's roofline.traffic)
"""
import argparse
import collections
import csv
import gzip
import io
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "smsp__cycles_elapsed.avg.per_second", "launch__cluster_dim_x",
]


def short_name(name):
    name = name.split("(")[0].strip()
    name = re.sub(r"^void\s+", "", name)
    return re.sub(r"^(?:[\w<>]+::)+", "", name)


def launch_list(src, rnd, out_dir):
    rows = []
    with open(src, newline="") as f:
        lines = f.readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    for r in csv.DictReader(io.StringIO("".join(lines[start:]))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((short_name(r["Kernel Name"]), v_us))
    forwards = sum(1 for n, _ in rows if n.startswith("sls_tail_kernel")) or 1
    agg = collections.OrderedDict()
    for n, v in rows:
        c = agg.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += v
    total = sum(c[1] for c in agg.values()) / forwards
    dst = os.path.join(out_dir, f"{rnd}_launches_bench_step.csv")
    with open(dst, "w") as f:
        f.write("kernel,launches_per_step,us_per_step,avg_us,share\n")
        other = [0, 0.0]
        for n, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if cnt % forwards != 0:                                 # not once-per-forward: weight packing, synthetic inputs, torch glue
                other[0] += cnt
                other[1] += us
                continue
            f.write(f'"{n}",{cnt / forwards:g},{us / forwards:.1f},{us / cnt:.1f},{us / forwards / total:.3f}\n')
        f.write(f'"(set-up: weight packing, synthetic clips, bench glue - not per forward)",{other[0] / forwards:.2f},'
                f"{other[1] / forwards:.1f},,{other[1] / forwards / total:.3f}\n")
        f.write(f'"TOTAL (sum of kernel durations over {forwards} forwards, ncu: cold caches, serialised)",'
                f"{len(rows) / forwards:g},{total:.1f},,1.000\n")
    with open(src, "rb") as fi, gzip.open(os.path.join(out_dir, f"{rnd}_launches_bench_raw.csv.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)
    print("wrote", dst, f"({forwards} forwards, {len(rows)} launches)")


def rep_summary(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True)
    if out.returncode != 0:
        print("ncu failed on", rep, out.stderr[-300:])
        return []
    rd = list(csv.reader(io.StringIO(out.stdout)))
    hdr_i = next(i for i, r in enumerate(rd) if r and r[0] == "ID")
    hdr, units, body = rd[hdr_i], rd[hdr_i + 1], rd[hdr_i + 2:]
    keep = ["Kernel Name", "Grid Size", "Block Size"] + [m for m in METRICS if m in hdr]
    idx = [hdr.index(k) for k in keep]
    recs = []
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([units[i] for i in idx])
        for r in body:
            if len(r) < len(hdr):
                continue
            w.writerow([short_name(r[i]) if k == "Kernel Name" else r[i] for k, i in zip(keep, idx)])
            recs.append({k: r[i] for k, i in zip(keep, idx)} | {"_units": {k: units[i] for k, i in zip(keep, idx)}})
    print("wrote", dst, f"({len(recs)} launches)")
    return recs


def dram_bytes(rec):
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        u = rec["_units"][k].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        tot += float(rec[k].replace(",", "")) * scale
    return int(tot)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r01")
    ap.add_argument("--src", default=os.path.join(ROOT, "gpurun_out"))
    a = ap.parse_args()
    out_dir = os.path.join(ROOT, "profiles")
    ll = os.path.join(a.src, "launches_bench.csv")
    if os.path.exists(ll):
        launch_list(ll, a.round, out_dir)
    traffic = {"_source": "ncu --set full --clock-control none captures (profiles/%s_ncu_*_summary.csv): dram__bytes_read.sum + "
                          "dram__bytes_write.sum per launch, bytes" % a.round}
    for rep, tag in (("prof_gemm", "gemm"), ("prof_ln2", "conv_pair"), ("prof_hbm", "hbm_kernels"), ("prof_pool", "sls_pool"),
                     ("prof_sae", "sae_head"), ("prof_win", "window_head")):
        path = os.path.join(a.src, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        recs = rep_summary(path, os.path.join(out_dir, f"{a.round}_ncu_{tag}_summary.csv"))
        by = collections.OrderedDict()
        for r in recs:
            by.setdefault(short_name(r["Kernel Name"]), []).append(dram_bytes(r))
        for n, v in by.items():
            traffic[n] = int(sum(v) / len(v))
            traffic[n + "_launches"] = v
    if len(traffic) > 1:
        # bench.py reads "tc_gemm_pair_kernel" (mean over one qkv / out_proj / fc1 / fc2 launch each)
        with open(os.path.join(out_dir, "traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
        print("wrote profiles/traffic.json")


if __name__ == "__main__":
    main()
