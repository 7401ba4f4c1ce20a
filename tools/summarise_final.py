"""Turns the outputs of tools/r02_final.sh (gpurun_out/<dir>) into the tracked evidence under profiles/:
    python tools/summarise_final.py gpurun_out/r02final2 r02c
bench lines, launch-list summary (+ the gzipped list), DRAM traffic of the DMMA launches (traffic.json), the ncu full
captures' key metrics, the overlapping-handle logs and the sanitizer / pytest tails."""
import collections
import csv
import gzip
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag = sys.argv[1], sys.argv[2]
P = os.path.join(ROOT, "profiles")


def last_json_line(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return lines[-1] if lines else None


for c in ("", "_cfg1", "_cfg3", "_cfg5"):
    f = os.path.join(src, f"bench{c}.json")
    if os.path.isfile(f) and last_json_line(f):
        open(os.path.join(P, f"{tag}_bench{c}.json"), "w").write(last_json_line(f) + "\n")

# ---- launch list: the second evaluation ----
lf = os.path.join(src, "launches.csv")
if os.path.isfile(lf):
    rows = list(csv.reader(open(lf)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = {x: i for i, x in enumerate(rows[hi])}
    data = [r for r in rows[hi + 1:] if len(r) > h["Metric Value"] and r[h["ID"]].isdigit()]
    starts = [i for i, r in enumerate(data) if "cov_build" in r[h["Kernel Name"]]]
    ev = data[starts[-1]:] if len(starts) > 1 else data
    agg = collections.OrderedDict()
    for r in ev:
        name = r[h["Kernel Name"]]
        short = name.split("(")[0]
        if "gemm_dmma_kernel" in name:
            short = ("void lk::gemm_dmma_kernel<0, 0>" if "(bool)0, (bool)0" in name or "<0, 0>" in name else
                     "void lk::gemm_dmma_kernel<0, 1>" if "(bool)0, (bool)1" in name or "<0, 1>" in name else
                     "void lk::gemm_dmma_kernel<1, 1>")
        v = float(r[h["Metric Value"]].replace(",", ""))
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[h["Metric Unit"]], 1e-6)
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{tag}_launches_summary.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, one LL+grad evaluation at n = 20000, d = 10 (the second "
                "of two; tools/r02_final.sh); cold-cache, serialised launches\nkernel,launches,time_ms,share\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{a[0]},{a[1]:.3f},{a[1] / tot:.4f}\n')
    with open(lf, "rb") as fi, gzip.open(os.path.join(P, f"{tag}_launches.csv.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)

# ---- DRAM traffic ----
tf = os.path.join(src, "gemm_traffic.csv")
if os.path.isfile(tf):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_traffic_json.py"), tf,
                    f"{tag}: band / serpentine tile tables, elected-lane producer, 3D boxes"], check=True)
    with open(tf, "rb") as fi, gzip.open(os.path.join(P, f"{tag}_gemm_traffic.csv.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)

# ---- full captures ----
keys = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__ops_path_tensor_src_fp64.sum.per_second",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "sm__cycles_elapsed.avg.per_second"]
with open(os.path.join(P, f"{tag}_gemm_dmma_ncu_full.csv"), "w") as out:
    out.write("# ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s {353 | 12 | 351 -c 2} -c 1 python "
              f"tools/profile_eval.py 20000 10 1 (tools/r02_final.sh, {src})\nlabel,metric,unit,value\n")
    for k in ("lauum", "syrk", "trtri"):
        rep = os.path.join(src, f"prof_{k}.ncu-rep")
        if not os.path.isfile(rep):
            continue
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(r.stdout.splitlines()))
        if len(rows) < 3:
            continue
        hh, uu = rows[0], rows[1]
        for ri, v in enumerate(rows[2:]):
            label = k if len(rows) == 3 else f"{k}_{ri}"
            for key in keys:
                if key in hh:
                    i = hh.index(key)
                    out.write(f"{label},{key},{uu[i]},{v[i]}\n")

# ---- logs ----
with open(os.path.join(P, f"{tag}_final_run.txt"), "w") as f:
    for name, n in (("gpu.txt", 5), ("pytest_gpu.log", 8), ("smoke.log", 2), ("concurrent_n20000.log", 4), ("concurrent_n5000.log", 8),
                    ("racecheck.log", 4), ("memcheck.log", 4)):
        fp = os.path.join(src, name)
        if os.path.isfile(fp):
            f.write(f"== {name}\n" + "\n".join(open(fp).read().splitlines()[-n:]) + "\n")
print("profiles updated from", src)
