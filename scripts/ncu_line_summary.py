"""Instructions executed and stall samples per CUDA source line of one kernel of an ncu report (needs -lineinfo
and --import-source on):  python scripts/ncu_line_summary.py file.ncu-rep [kernel-regex] [top-n]"""
import collections, csv, subprocess, sys
cmd = ["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 2 and sys.argv[2]:
    cmd += ["--kernel-name", "regex:" + sys.argv[2]]
rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
inst, smp, text = collections.Counter(), collections.Counter(), {}
fname, h, line = "", None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        h = r
        IE, SM = h.index("Instructions Executed"), h.index("# Samples")
    elif h and len(r) > IE:
        if r[0]:
            line = (fname, int(r[0]))
            text[line] = r[1].strip()[:90]
        if r[2] and line:
            num = lambda x: int(x) if x.isdigit() else 0
            inst[line] += num(r[IE])
            smp[line] += num(r[SM])
tot, ts = sum(inst.values()), sum(smp.values())
print(f"total warp-instructions {tot}, samples {ts}")
for k, n in inst.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30):
    print(f"{k[0]}:{k[1]:<5d} {100.0 * n / tot:5.1f}% inst {100.0 * smp[k] / max(ts, 1):5.1f}% smp  {text.get(k, '')}")
