"""Per-source-line roll-up of an ncu report's source page (needs -lineinfo): samples, executed warp instructions and
SASS size per line.  usage: python scripts/ncu_lines.py report.ncu-rep [kernel-regex] [top]"""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else "."; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = func = None; hdr = None; cur = None; seen = set()
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = r[1]; hdr = None; continue
    if r[0] == "Line No": hdr = r; ci = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or not re.search(pat, func or ""): continue
    if (func, "k") not in seen and r[0] != "": pass
    if r[0] != "":
        cur = (fname, int(r[0]), r[1].strip())
        agg.setdefault(cur, [0.0, 0.0, 0])
    elif cur is not None:
        f = lambda x: float(x) if x not in ("", "-") else 0.0
        a = agg[cur]; a[0] += f(r[ci["# Samples"]]); a[1] += f(r[ci["Instructions Executed"]]); a[2] += 1
ts = sum(a[0] for a in agg.values()) or 1; ti = sum(a[1] for a in agg.values()) or 1; tn = sum(a[2] for a in agg.values())
print(f"total samples {ts:.0f}, executed warp instructions {ti:.0f}, SASS instructions {tn} (all captured launches of the kernel summed)")
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:<5d} {100*a[0]/ts:5.1f}% samples {100*a[1]/ti:5.1f}% inst {a[2]:5d} sass  {src[:100]}")
