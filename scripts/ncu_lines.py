#!/usr/bin/env python3
"""Summarise an .ncu-rep per CUDA source line: share of issued instructions, lane utilisation, top stalls.
usage: ncu_lines.py report.ncu-rep [kernel-substring] [topN]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; ksub = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None; agg = collections.OrderedDict(); kernel = None; seen_kernels = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        kernel = r[1]
        if kernel not in seen_kernels: seen_kernels.append(kernel)
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "": continue
    if ksub and ksub not in (kernel or ""): continue
    ci = {n: i for i, n in enumerate(hdr)}
    def g(n):
        try: return float(r[ci[n]])
        except Exception: return 0.0
    key = (cur_file, r[0], r[1].strip()[:90])
    a = agg.setdefault(key, collections.Counter())
    a["inst"] += g("Instructions Executed"); a["thr"] += g("Thread Instructions Executed"); a["samples"] += g("# Samples")
    for n in hdr:
        if n.startswith("stall_") and "Not Issued" not in n: a[n] += g(n)
tot = sum(a["inst"] for a in agg.values()); tott = sum(a["thr"] for a in agg.values()); tots = sum(a["samples"] for a in agg.values())
print("kernels:", len(seen_kernels), "| warp-instr %.3g  thread-instr %.3g  avg lanes %.2f" % (tot, tott, tott / max(tot, 1)))
for (f, ln, src), a in sorted(agg.items(), key=lambda x: -x[1]["inst"])[:top]:
    st = sorted(((n, v) for n, v in a.items() if n.startswith("stall_")), key=lambda x: -x[1])[:2]
    print("%5.2f%% inst %5.2f%% smp lanes %4.1f  %-16s:%-4s %s   [%s]" % (100 * a["inst"] / tot, 100 * a["samples"] / max(tots, 1), a["thr"] / max(a["inst"], 1), f, ln, src,
          ", ".join("%s %.0f%%" % (n[6:], 100 * v / max(a["samples"], 1)) for n, v in st)))
