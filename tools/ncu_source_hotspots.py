"""Per-CUDA-source-line stall attribution from an ncu report.
usage: python tools/ncu_source_hotspots.py <report.ncu-rep> <kernel substring> [top]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
def I(v):
    try: return int(v)
    except Exception: return 0
res = {}
fp = fn = hdr = None
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": fp = r[1]; hdr = None; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit() or pat not in fn: continue
    key = (fp.split('/')[-1], int(r[0]), r[1].strip()[:90])
    e = res.setdefault(key, {"samples": 0, "inst": 0, "st": {}})
    e["samples"] += I(r[hdr.index('# Samples')]); e["inst"] += I(r[hdr.index('Instructions Executed')])
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            e["st"][h] = e["st"].get(h, 0) + I(r[i])
tot = sum(e["samples"] for e in res.values()) or 1
toti = sum(e["inst"] for e in res.values()) or 1
print("total samples", tot, "warp instructions", toti)
for k, e in sorted(res.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(e["st"].items(), key=lambda kv: -kv[1])[:3]
    print(f"{e['samples']/tot*100:5.1f}% smp {e['inst']/toti*100:5.1f}% inst  {k[0]}:{k[1]:4d}  {k[2][:70]:70s} {[(a[6:], b) for a, b in st if b]}")
