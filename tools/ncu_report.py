"""Summarise every kernel of an .ncu-rep into one small markdown file (run on the GPU box; the report itself is too big to bring back):
   python tools/ncu_report.py report.ncu-rep out.md [lines_per_kernel]"""
import csv, subprocess, sys, io, collections

rep, out = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 22
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
h, units, data = raw[0], raw[1], raw[2:]
ki = h.index("Kernel Name")
seen = collections.OrderedDict()
for r in data:
    name = r[ki].split("(")[0].split("::")[-1]
    seen.setdefault(name, r)          # first captured launch of each kernel
w = open(out, "w")
w.write("# ncu --set full summary of %s (first captured launch of each kernel)\n\n" % rep.split("/")[-1])
w.write("| kernel | " + " | ".join(k.replace("launch__", "").replace(".sum", "").replace(".avg.pct_of_peak_sustained", "%") for k in KEYS) + " |\n")
w.write("|---|" + "---|" * len(KEYS) + "\n")
for name, r in seen.items():
    vals = []
    for k in KEYS:
        if k in h:
            i = h.index(k); vals.append("%s %s" % (r[i], units[i]))
        else:
            vals.append("-")
    w.write("| %s | %s |\n" % (name, " | ".join(vals)))
w.write("\n")

for name in seen:
    txt = run(["--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:^%s$" % name, "-c", "1"]) if False else \
          run(["--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:%s" % name.replace("void ", "").split("<")[0], "--launch-count", "1"])
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = None; cur_file = ""; lines = collections.OrderedDict()
    for r in rows:
        if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
        if len(r) > 6 and r[0] == "Line No": hdr = r; continue
        if hdr is None or len(r) != len(hdr) or r[0] == "": continue
        key = (cur_file, int(r[0]))
        d = lines.setdefault(key, dict(src=r[1], samples=0, inst=0, stalls=collections.Counter()))
        d["samples"] += int(r[hdr.index("# Samples")] or 0)
        d["inst"] += int(r[hdr.index("Instructions Executed")] or 0)
        for i, hh in enumerate(hdr):
            if hh.startswith("stall_") and "Not Issued" not in hh and r[i] not in ("", "0"):
                d["stalls"][hh[6:]] += int(r[i])
    ts = sum(d["samples"] for d in lines.values()) or 1; ti = sum(d["inst"] for d in lines.values()) or 1
    w.write("## %s -- hot lines (samples %d, warp instructions %d)\n\n| file:line | samples | inst | top stalls | source |\n|---|---|---|---|---|\n" % (name, ts, ti))
    for (f, ln), d in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:topn]:
        st = " ".join("%s:%d%%" % (k, 100 * v // max(1, d["samples"])) for k, v in d["stalls"].most_common(3))
        w.write("| %s:%d | %.1f%% | %.1f%% | %s | `%s` |\n" % (f, ln, 100 * d["samples"] / ts, 100 * d["inst"] / ti, st, d["src"].strip()[:110].replace("|", "\\|")))
    w.write("\n")
w.close()
print("wrote", out)
