"""Summarise an ncu report per CUDA source line: python tools/ncu_lines.py rep.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H = None
cur = None
agg = collections.OrderedDict()
for r in rows:
    if "# Samples" in r:
        H = r
        continue
    if not H or len(r) != len(H):
        continue
    ni, ii = H.index("# Samples"), H.index("Instructions Executed")
    if r[0] not in ("", "-") and r[2] in ("-", ""):   # a CUDA line row
        cur = (r[0], r[1][:100])
        agg.setdefault(cur, [0, 0])
    elif cur is not None:  # SASS row under the current CUDA line
        try:
            agg[cur][0] += int(r[ni] or 0); agg[cur][1] += int(r[ii] or 0)
        except ValueError:
            pass
tot_s = sum(v[0] for v in agg.values()); tot_i = sum(v[1] for v in agg.values())
print(f"samples {tot_s} warp-instructions {tot_i}")
for (ln, src), (s_, i_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{i_:9d} ({100*i_/max(tot_i,1):4.1f}%) smp {s_:4d}  L{ln:>4}: {src}")
