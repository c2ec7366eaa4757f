"""Developer tool: summarise an `ncu --set full` report into the two artefacts committed under profiles/:
a per-kernel CSV of the metrics DESIGN.md quotes and the DRAM-traffic table bench.py reads.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_full_final.csv profiles/r01_traffic.json
"""
import csv, io, json, subprocess, sys

rep, out_csv, out_json = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_tensor_op_imma.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__warps_eligible.avg.per_cycle_active"]
cols = [w for w in want if w in hdr]
ix = {w: hdr.index(w) for w in cols}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(unit, 1)


traffic = {}
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([units[ix[c]] for c in cols])
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        w.writerow([r[ix[c]] for c in cols])
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        t = traffic.setdefault(name, {"dram_bytes_per_launch": 0.0, "ncu_duration_us": 0.0, "launches": 0, "warp_instructions": 0.0})
        t["dram_bytes_per_launch"] += rd + wr
        t["ncu_duration_us"] += to_us(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
        t["warp_instructions"] += float(r[ix["smsp__inst_executed.sum"]].replace(",", ""))
        t["launches"] += 1
for t in traffic.values():
    for k in ("dram_bytes_per_launch", "ncu_duration_us", "warp_instructions"):
        t[k] /= t["launches"]
json.dump(traffic, open(out_json, "w"), indent=1)
print(json.dumps(traffic, indent=1))
