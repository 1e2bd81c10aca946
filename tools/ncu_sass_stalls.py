"""Per-SASS-instruction view of an ncu report: executed counts, stall samples and the dominant stall reasons.
usage: python tools/ncu_sass_stalls.py <report.ncu-rep> [min_samples]   (first kernel in the report)"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; mins = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = []; hdr = None; k = 0
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        k += 1
        if k == 2: break
        continue
    if r and r[0] == "Address": hdr = r; continue
    if r and r[0].startswith("0x"): rows.append(r)
ie = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); ismp = hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ie]) for r in rows); ts = sum(int(r[ismp]) for r in rows)
agg = collections.Counter()
for r in rows:
    for i in stall:
        try: agg[hdr[i][6:]] += int(r[i])
        except ValueError: pass
print("warp instructions", tot, "samples", ts, dict(agg.most_common(10)))
cls = collections.Counter(); smp = collections.Counter()
for r in rows: cls[int(r[ie])] += 1; smp[int(r[ie])] += int(r[ismp])
print("exec-count classes (count, #sass, samples):", [(c, n, smp[c]) for c, n in sorted(cls.items(), key=lambda kv: -kv[0] * kv[1])[:8]])
for i, r in enumerate(rows):
    if int(r[ismp]) >= mins:
        st = sorted(((hdr[j][6:], int(r[j])) for j in stall if r[j].isdigit() and int(r[j]) > 0), key=lambda kv: -kv[1])[:3]
        prev = rows[i - 1][isrc].strip()[:40] if i else ""
        print(f"{i:5d} x{int(r[ie]):8d} {int(r[ismp]):5d} {r[isrc].strip()[:60]:60s} {st}   <- {prev}")
