#!/usr/bin/env python3
"""Share of warp / thread instructions per STAGE of k_pool from an .ncu-rep captured with --import-source on.
Stages are found by marker strings in the CURRENT sources (a line belongs to the last marker above it), so the table
follows the code.  usage: ncu_stages.py report.ncu-rep"""
import csv, subprocess, io, collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARK = {
    "sched_pool.cuh": [("template <bool WIDE", "pool: prologue"), ("// ---- (a) lanes without a ray", "pool: (a) take ready slots + loop head"),
                       ("// ---- (b) management section", "pool: (b) slot -> registers"), ("// ---- (1)-(4)", "pool: chain_event glue"),
                       ("// ---- (5) end of the sample", "pool: (5) end of sample"), ("// ---- (6) fetch chains", "pool: (6) fetch chains"),
                       ("// ---- (7) next camera ray", "pool: (7) camera ray"), ("// ---- (8) set the ray up", "pool: (8) ray set-up + registers -> slot, queue push"),
                       ("// ---- (c) one traversal quantum", "pool: (c) traversal quantum glue"), ("warp_add(&s.stats[ST_RADIANCE]", "pool: epilogue")],
    "estimator.cuh": [("__device__ __forceinline__ bool hits_emitter_bounds", "est: hits_emitter_bounds"), ("__device__ __forceinline__ bool trace_one", "est: diagnostics"),
                      ("__device__ __forceinline__ float3 shading_normal", "est: shading normal / camera ray / seed"), ("__device__ __forceinline__ void emitter_cone", "est: emitter_cone"),
                      ("// ---- (1) material dispatch", "est: (1) material dispatch"), ("// ---- (2) retire the finished shadow ray", "est: (2) retire shadow ray"),
                      ("// ---- (3) shoot_ray_to_light", "est: (3) per-job filter set-up"), ("const unsigned jobs = __ballot_sync", "est: (3) cooperative tries loop"),
                      ("// the survivors are confirmed", "est: (3) confirmation of survivors"), ("// ---- (4) end of the opaque branch", "est: (4) light term + bounce")],
    "traverse.cuh": [("__device__ __forceinline__ RayPre ray_precompute", "trav: triangle test"), ("struct TravStack", "trav: stack"),
                     ("__device__ __forceinline__ float3 safe_rcp_dir", "trav: safe_rcp_dir (ray set-up, emitter bounds)"), ("struct StepRay", "trav: step_ray"),
                     ("__device__ __forceinline__ bool step_tri(", "trav: triangle fetch"), ("static __constant__ uint32_t c_unit_magic", "trav: wide node step (incl. byte_to_unit)"),
                     ("// ---- binary ---", "trav: binary node step")],
    "common.cuh": [("namespace lisa {", "common: float3 helpers (confirmation, bounce, shading: low lane counts)"), ("// ---- RNG", "common: rng / shoot_ray_hemisphere / fresnel / refract"),
                   ("// ---- tonemap", "common: misc")],
}
table = {}
for f, marks in MARK.items():
    lines = open(os.path.join(ROOT, "lisa_b200", "csrc", f)).read().splitlines()
    pts = []
    for m, label in marks:
        hit = [i + 1 for i, l in enumerate(lines) if m in l]
        if hit:
            pts.append((hit[0], label))
    table[f] = sorted(pts)

def bucket(f, ln):
    if f not in table:
        return f
    lab = f + ": head"
    for start, label in table[f]:
        if ln >= start:
            lab = label
    return lab

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; agg = collections.Counter(); aggw = collections.Counter()
def num(x):
    try: return float(x)
    except Exception: return 0.0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "" or not r[0].isdigit(): continue
    ci = {n: i for i, n in enumerate(hdr)}
    ln = int(r[0]); thr = num(r[ci["Thread Instructions Executed"]]); w = num(r[ci["Instructions Executed"]])
    k = bucket(cur, ln); agg[k] += thr; aggw[k] += w
T = sum(agg.values()); W = sum(aggw.values())
print("# total: %.4g warp instructions, %.4g thread instructions, %.2f lanes per instruction" % (W, T, T / W))
for k, v in sorted(aggw.items(), key=lambda x: -x[1]):
    print("%5.1f%% warp %5.1f%% thr lanes %4.1f  %s" % (100 * v / W, 100 * agg[k] / T, agg[k] / max(v, 1), k))
