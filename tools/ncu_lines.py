"""hot source lines of one kernel from an .ncu-rep (needs -lineinfo + --import-source on):
   python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io, collections

rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; cur_file = ""; lines = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0] == "": continue                 # SASS rows
    key = (cur_file, int(r[0]))
    d = lines.setdefault(key, dict(src=r[1], samples=0, inst=0, stalls=collections.Counter(), excess=0))
    d["samples"] += int(r[hdr.index("# Samples")] or 0)
    d["inst"] += int(r[hdr.index("Instructions Executed")] or 0)
    d["excess"] += int(r[hdr.index("L1 Wavefronts Shared Excessive")] or 0)
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h and r[i] not in ("", "0"):
            d["stalls"][h[6:]] += int(r[i])
ts = sum(d["samples"] for d in lines.values()) or 1; ti = sum(d["inst"] for d in lines.values()) or 1
print("total samples %d, warp instructions %d" % (ts, ti))
print("| file:line | samples | inst | top stalls | source |\n|---|---|---|---|---|")
for (f, ln), d in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:topn]:
    st = " ".join("%s:%d%%" % (k, 100 * v // max(1, d["samples"])) for k, v in d["stalls"].most_common(3))
    print("| %s:%d | %.1f%% | %.1f%% | %s | `%s` |" % (f, ln, 100 * d["samples"] / ts, 100 * d["inst"] / ti, st, d["src"].strip()[:100]))
