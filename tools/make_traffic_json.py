"""profiles/traffic.json from an ncu pass over the DMMA GEMM launches of one evaluation:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_dmma
        --clock-control none -c 354 --csv --log-file <csv> python tools/profile_eval.py 20000 10 1
    python tools/make_traffic_json.py <csv> <label> [out.json]
Keeps the previous file's per-kernel numbers under "history"."""
import csv
import json
import os
import sys

src, label = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = {x: i for i, x in enumerate(rows[hi])}
per = {}
ids = set()
for r in rows[hi + 1:]:
    if len(r) <= h["Metric Value"] or not r[h["ID"]].isdigit():
        continue
    name = r[h["Kernel Name"]]
    key = "gemm_dmma_kernel<0, 0>" if "<0, 0>" in name or "(bool)0, (bool)0" in name else \
          "gemm_dmma_kernel<0, 1>" if "<0, 1>" in name or "(bool)0, (bool)1" in name else "gemm_dmma_kernel<1, 1>"
    d = per.setdefault(key, {"launches": set(), "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "ms_under_ncu": 0.0})
    d["launches"].add(r[h["ID"]])
    ids.add(r[h["ID"]])
    v = float(r[h["Metric Value"]].replace(",", ""))
    unit = r[h["Metric Unit"]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3,
             "nsecond": 1e-6, "msecond": 1.0}.get(unit, 1.0)
    m = r[h["Metric Name"]]
    if m == "dram__bytes_read.sum":
        d["dram_read_bytes"] += v * scale
    elif m == "dram__bytes_write.sum":
        d["dram_write_bytes"] += v * scale
    elif m == "gpu__time_duration.sum":
        d["ms_under_ncu"] += v * scale
for d in per.values():
    d["launches"] = len(d["launches"])
rd = sum(d["dram_read_bytes"] for d in per.values())
wr = sum(d["dram_write_bytes"] for d in per.values())
old = json.load(open(out)) if os.path.isfile(out) else {}
hist = old.get("history", {})
if old.get("per_kernel"):
    hist[old.get("label", "r01c")] = {"gemm_dmma_kernel_bytes_per_eval": old.get("gemm_dmma_kernel_bytes_per_eval"),
                                      "per_kernel": old["per_kernel"]}
res = {"label": label, "gemm_dmma_kernel_bytes_per_eval": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
       "launches": len(ids),
       "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm_dmma, all DMMA launches of one LL+grad "
                 f"evaluation at n=20000 d=10 ({label}; tools/make_traffic_json.py)",
       "per_kernel": per, "algorithmic_bytes_per_eval": 16.0 * 20000 * 20000, "history": hist}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: res[k] for k in ("gemm_dmma_kernel_bytes_per_eval", "dram_read_bytes", "dram_write_bytes", "launches")}))
for k, d in per.items():
    print(k, d)
