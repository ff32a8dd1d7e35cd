#!/usr/bin/env python3
"""Per-region share of thread instructions for k_shadow / k_extend from an .ncu-rep (regions = source line ranges).
usage: ncu_regions.py report.ncu-rep"""
import csv, subprocess, io, collections, sys, re
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
src = {}
cur=None; hdr=None; agg=collections.Counter(); aggw=collections.Counter()
def num(x):
    try: return float(x)
    except Exception: return 0.0
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if hdr is None or r[0]=="" or not r[0].isdigit(): continue
    ci={n:i for i,n in enumerate(hdr)}
    ln=int(r[0]); thr=num(r[ci["Thread Instructions Executed"]]); w=num(r[ci["Instructions Executed"]])
    text=r[1]
    k=cur
    if cur in ("traverse.cuh","wavefront.cu","common.cuh","pool.cuh","estimator.cuh","estimator.cu","sched_pool.cuh","sched_path.cuh","sched_wavefront.cuh"):
        k="%s:%d-%d"%(cur, (ln//10)*10, (ln//10)*10+9)
    agg[k]+=thr; aggw[k]+=w
T=sum(agg.values()); W=sum(aggw.values())
print("total thread-instr %.3g warp-instr %.3g lanes %.1f"%(T,W,T/W))
for k,v in agg.most_common(45): print("%5.1f%% thr  %5.1f%% warp  lanes %4.1f  %s"%(100*v/T,100*aggw[k]/W, v/max(aggw[k],1), k))
