"""Developer tool: per-kernel figures of one tracking frame from two ncu outputs, as the JSON bench.py reads
(profiles/r02_kernel_profile.json):

    python tools/ncu_profile_json.py <launches.csv from --metrics gpu__time_duration.sum,smsp__inst_executed.sum> <full.ncu-rep or ''> <out.json> [first_frame_launch last_frame_launch]

launches.csv: every launch of tools/profile_frames.py; the kernels between two consecutive ingest_kernel launches are one frame.
"""
import csv
import io
import json
import subprocess
import sys

launch_csv, rep, out = sys.argv[1:4]
rows = [r for r in csv.reader(open(launch_csv)) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
idi = hdr.index("ID")
launches = {}
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    L = launches.setdefault(int(r[idi]), {"name": r[ki].split("(")[0].replace("void ", "").split("<")[0]})
    v = float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum":
        L["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1.0)
    elif r[mi] == "smsp__inst_executed.sum":
        L["inst"] = v
order = [launches[k] for k in sorted(launches)]
# frames = runs that start with an ingest kernel; take tracking frames only (no detection kernels inside)
starts = [i for i, L in enumerate(order) if L["name"].startswith("ingest_kernel")]
frames = []
for a, b in zip(starts, starts[1:]):
    fr = order[a:b]
    if any(L["name"].startswith(("fast_score", "cell_select", "klt_template", "lk_scharr")) for L in fr):
        continue
    if any(L["name"].startswith("sparse_align") for L in fr):
        frames.append(fr)
frames = frames[5:]          # steady state
kern = {}
for fr in frames:
    for L in fr:
        k = kern.setdefault(L["name"], {"launches": 0, "us": 0.0, "warp_instructions": 0.0})
        k["launches"] += 1
        k["us"] += L.get("us", 0.0)
        k["warp_instructions"] += L.get("inst", 0.0)
nf = len(frames)
res = {"source": "ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none on tools/profile_frames.py (one C3 sequence, "
                 "kernel-by-kernel launches; cold-cache, serialised: compare shares, not absolutes)",
       "frames_averaged": nf, "kernels": {}}
tot_us = sum(k["us"] for k in kern.values()) / nf
tot_inst = sum(k["warp_instructions"] for k in kern.values()) / nf
for name, k in sorted(kern.items(), key=lambda kv: -kv[1]["us"]):
    res["kernels"][name] = {"launches_per_frame": k["launches"] / nf, "ncu_us_per_frame": k["us"] / nf, "share_of_frame": k["us"] / nf / tot_us,
                            "warp_instructions_per_frame": k["warp_instructions"] / nf}
res["ncu_us_per_frame"] = tot_us
res["warp_instructions_per_frame"] = tot_inst
if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    h, u = rr[0], rr[1]

    def col(n):
        return h.index(n) if n in h else None
    want = {"dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum", "registers": "launch__registers_per_thread",
            "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
            "duration": "gpu__time_duration.sum", "grid": "launch__grid_size", "block": "launch__block_size"}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rr[2:]:
        if len(r) < len(h):
            continue
        name = r[col("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0]
        a = acc.setdefault(name, {"n": 0})
        a["n"] += 1
        for key, m in want.items():
            c = col(m)
            if c is None:
                continue
            v = float(r[c].replace(",", ""))
            if key.startswith("dram"):
                v *= scale.get(u[c], 1)
            a[key] = a.get(key, 0.0) + v
    for name, a in acc.items():
        rec = res["kernels"].setdefault(name, {})
        n = a.pop("n")
        for key, v in a.items():
            rec["full_" + key] = v / n
        if "dram_read" in a:
            rec["dram_bytes_per_launch"] = (a["dram_read"] + a["dram_write"]) / n
        rec["full_set_launches_captured"] = n
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1)[:3000])
