#!/usr/bin/env python3
"""Share of warp / thread instructions per STAGE of k_pool (line ranges of sched_pool.cuh and estimator.cuh) from an
.ncu-rep captured with --import-source on.  usage: ncu_stages.py report.ncu-rep"""
import csv, subprocess, io, collections, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur=None; hdr=None; agg=collections.Counter(); aggw=collections.Counter()
def num(x):
    try: return float(x)
    except Exception: return 0.0
def bucket(f, ln):
    if f=='sched_pool.cuh':
        if ln<77: return 'pool: prologue'
        if ln<=99: return 'pool: (a) take ready slots + loop head'
        if ln<=122: return 'pool: (b) slot->regs'
        if ln<=135: return 'pool: chain_event glue'
        if ln<=147: return 'pool: (5) end sample'
        if ln<=171: return 'pool: (6) fetch chains'
        if ln<=180: return 'pool: (7) camera ray'
        if ln<=216: return 'pool: (8) ray setup + regs->slot'
        if ln<=225: return 'pool: queue push'
        if ln<=293: return 'pool: (c) traversal quantum glue'
        return 'pool: epilogue'
    if f=='estimator.cuh':
        if ln<100: return 'est: hits_emitter_bounds'
        if ln<165: return 'est: shading normal/camera/seed'
        if ln<=185: return 'est: emitter_cone'
        if ln<=262: return 'est: (1) material dispatch'
        if ln<=273: return 'est: (2) retire shadow'
        if ln<=308: return 'est: (3) tries setup'
        if ln<=331: return 'est: (3) coop tries loop'
        if ln<=350: return 'est: (3) confirm loop'
        return 'est: (4) finish/bounce'
    if f=='traverse.cuh':
        if ln<=95: return 'trav: triangle test'
        if ln<=125: return 'trav: stack'
        if ln<=135: return 'trav: safe_rcp_dir (ray set-up, emitter bounds)'
        if ln<=195: return 'trav: step_ray, extract_byte'
        if ln<=212: return 'trav: triangle fetch'
        return 'trav: wide node step (incl. byte_to_unit)'
    if f=='common.cuh':
        if ln<=45: return 'common: float3 helpers (confirm loop, bounce, shading: low lane counts)'
        return 'common: rng / shoot_ray_hemisphere (confirm loop, bounce)'
    return f
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if hdr is None or r[0]=="" or not r[0].isdigit(): continue
    ci={n:i for i,n in enumerate(hdr)}
    ln=int(r[0]); thr=num(r[ci["Thread Instructions Executed"]]); w=num(r[ci["Instructions Executed"]])
    k=bucket(cur,ln); agg[k]+=thr; aggw[k]+=w
T=sum(agg.values()); W=sum(aggw.values())
for k,v in sorted(aggw.items(), key=lambda x:-x[1]): print("%5.1f%% warp %5.1f%% thr lanes %4.1f  %s"%(100*v/W,100*agg[k]/T, agg[k]/max(v,1), k))
