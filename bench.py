#!/usr/bin/env python3
"""bench.py — headline benchmark of the render path (BASELINE.json: Msamples/s, Cornell 2000x2000, 7 bounces).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--spp-per-step S]

Workload (config.workload): BASELINE.json configs[1] — the README Cornell box (scenes/cornell_c2.rto:
2000x2000, 7 bounces, 1002 authored triangles).  A STEP is one subframe of S samples per pixel over the
full image (S = 50 by default: the 2000-spp render is 40 such steps; subframes are the reference's own
unit of independently seeded samples, shader.cu:140-141, render.cc:75-131).  Every step renders a NEW
subframe index, so no step can reuse an earlier result.

  value   Msamples/s with the scene and its BVH resident in HBM (device time, CUDA events).
  e2e     the same metric through the C ABI with HOST buffers every step: lisa_create (H2D of the soup +
          device BVH build) + lisa_render_subframes + lisa_read_accum (D2H of the float4 image) + destroy.
  roofline  dominant kernel k_pool (the whole estimator, one persistent launch per step): algorithmic bytes per
          launch / its launch time (CUDA events around the launch, LISA_FLAG_PROFILE_STAGES) against the
          measured HBM peak, plus the issue-slot figures of the committed ncu capture (the kernel is issue bound).
  --pipeline path|pool|wavefront  forces one schedule of the estimator (ablation; same images).  Default: the
          library's own choice per tile, which is k_pool for this workload.
  cpu_baseline  the oracle (oracle/cpu_ref.c, OpenMP) on a bounded pixel sample of the same workload.
  --impl reference  the UNMODIFIED reference (gaetanserre/LiSA OptiX renderer built headless from its own
          sources, oracle/_ref/lisa_optix_ref) on the same config; falls back to the oracle port if OptiX
          cannot start.  The reference has no CPU renderer.
N > 1 (torchrun, one rank per GPU): sample-space partition — every rank renders its own subframe of the
full image each step, then ONE reduce of the float4 sums onto rank 0 (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.chdir(ROOT)

SCENE = "scenes/cornell_c2.rto"
METRIC = "Msamples/s (Cornell 2000x2000, 7 bounces)"


_JSON_FD = None


def guard_stdout():
    """From here on fd 1 is stderr for everybody (the OBJ loader's progress lines, NCCL's `NCCL version ...` banner when the
    box sets NCCL_DEBUG): the ONE JSON line goes to the real stdout through emit()."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def parse_quiet(fe, path):
    """fe.parse_scene with the loader's `Importing ...` lines (the reference prints them on stdout, parse_obj.cc) sent to
    stderr, so that this script's stdout is the ONE JSON line."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    try:
        return fe.parse_scene(path)
    finally:
        sys.stdout.flush()
        import ctypes
        ctypes.CDLL(None).fflush(None)  # the loader writes through C stdio: flush its buffer while fd 1 still points at stderr
        os.dup2(keep, 1)
        os.close(keep)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(pw) if pw else None}


def cpu_baseline(sc, spp, target_s=12.0):
    """Oracle port on a bounded random pixel sample of the workload (same seeds, same estimator)."""
    import numpy as np
    from oracle import binding
    S = binding.Scene(sc["vertices"], sc["normals"], sc["mat_indices"], sc["materials_packed"])
    cam = sc["camera"]
    rng = np.random.default_rng(0)
    w, h = sc["width"], sc["height"]
    probe = rng.choice(w * h, size=512, replace=False).astype(np.uint32)
    t0 = time.perf_counter()
    _, cnt = S.render(cam["eye"], cam["look_at"], cam["fov"], w, h, sc["num_bounces"], spp, pixels=probe)
    dt = time.perf_counter() - t0
    n = int(min(w * h, max(1024, 512 * target_s / max(dt, 1e-3))))
    pix = rng.choice(w * h, size=n, replace=False).astype(np.uint32)
    t0 = time.perf_counter()
    _, cnt = S.render(cam["eye"], cam["look_at"], cam["fov"], w, h, sc["num_bounces"], spp, pixels=pix)
    dt = time.perf_counter() - t0
    return {"value": round(cnt["samples"] / dt / 1e6, 5), "unit": "Msamples/s", "cores": cnt["threads"], "kind": "port",
            "sample": "%d random pixels of the %dx%d image x %d spp (%.1f s, %.1f rays/sample), oracle/cpu_ref.c + OpenMP"
                      % (n, w, h, spp, dt, (cnt["radiance_rays"] + cnt["shadow_rays"]) / max(cnt["samples"], 1))}


def run_reference(args, rank):
    """Reference arm: the unmodified reference's OptiX renderer (it has no CPU implementation)."""
    if rank != 0:
        return 0
    import lisa_b200.frontend as fe
    sc = fe.parse_scene(SCENE, load_meshes=False)
    w, h, S, K, W = sc["width"], sc["height"], args.spp_per_step, args.steps, args.warmup
    exe = os.path.join(ROOT, "oracle", "_ref", "lisa_optix_ref")
    base = {"metric": METRIC, "unit": "Msamples/s", "n_gpus": 1, "steps": K, "warmup": W, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "BASELINE configs[1]: README Cornell box 2000x2000, 7 bounces; step = one subframe of %d spp" % S,
                       "width": w, "height": h, "bounces": sc["num_bounces"], "spp_per_step": S}}
    out = None
    if os.path.exists(exe):
        try:
            os.makedirs("out", exist_ok=True)
            clk = ClockSampler(0)
            r = subprocess.run([exe, "-s", SCENE, "--spp", str(S), "--subframes", str(K), "--warmup-steps", str(W)],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=3000)
            clocks = clk.stop()
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode == 0 and line:
                j = json.loads(line[-1])
                ms = sum(j["step_ms"])
                v = w * h * S * K / ms / 1e3
                out = dict(base, value=round(v, 4), ms_per_step=round(ms / K, 3), clocks=clocks, gpu_launches=K,
                           cpu_baseline={"value": round(v, 4), "unit": "Msamples/s", "cores": 1, "kind": "reference",
                                         "sample": "unmodified gaetanserre/LiSA OptiX 7.4 renderer (oracle/_ref/lisa_optix_ref) on "
                                                   "1 B200 (software traversal, no RT cores), 1 host thread; %d launches of %d spp, "
                                                   "std::chrono around optixLaunch+sync" % (K, S)},
                           e2e={"value": round(v, 4), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
            else:
                sys.stderr.write("reference OptiX harness failed (rc %d): %s\n" % (r.returncode, r.stderr[-400:]))
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("reference OptiX harness failed: %r\n" % (e,))
    if out is None:
        # OptiX could not start: time the CPU restatement instead (labelled as a port)
        sc = parse_quiet(fe, SCENE)
        cb = cpu_baseline(sc, S, target_s=float(os.environ.get("LISA_BENCH_CPU_TARGET_S", "20")))  # env: the CPU tests shorten it
        out = dict(base, value=cb["value"], ms_per_step=round(w * h * S / (cb["value"] * 1e3), 1), cpu_baseline=cb, gpu_launches=0,
                   clocks={"sm_mhz": None, "sm_max_mhz": None, "reasons": []},
                   e2e={"value": cb["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    emit(out)
    return 0


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--pipeline", default="auto", choices=["auto", "pool", "path", "wavefront"])
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import numpy as np
    import torch
    import torch.distributed as dist
    import lisa_b200.frontend as fe
    import lisa_b200.rt as rt
    from lisa_b200 import dist as ldist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sc = parse_quiet(fe, SCENE)
    w, h, S, K, W = sc["width"], sc["height"], args.spp_per_step, args.steps, max(args.warmup, 0)
    npix = w * h
    pipe_flag = rt.FLAG_WAVEFRONT if args.pipeline == "wavefront" else 0
    if args.pipeline != "auto":
        os.environ["LISA_PIPELINE"] = args.pipeline   # read by lisa_create (ablation switch)
    R = rt.Renderer.from_scene(sc, device=local_rank, flags=rt.FLAG_PROFILE_STAGES | pipe_flag)
    # L2 is flushed between steps (a buffer twice its size is rewritten): the kernel's own inputs (BVH, triangles:
    # ~150 KB) are far smaller than L2, so nothing may survive from the previous step
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    acc_t = ldist.accum_tensor(R) if world > 1 else None
    total = torch.zeros_like(acc_t) if (world > 1 and rank == 0) else None

    def step(i):
        """One pass of the hot path: this rank's subframe of step i (+ the one reduce when N > 1)."""
        R.reset()
        flush.fill_(i & 0xff)
        torch.cuda.current_stream().synchronize()
        R.render_subframes(i * world + rank, 1, S)
        if world > 1:
            dist.reduce(acc_t, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                total.add_(acc_t)
        return R.stats()

    for i in range(W):
        step(i)
    barrier()
    clk = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agg = dict(render_ms=0.0, shadow_ms=0.0, extend_ms=0.0, shadow_launches=0, extend_launches=0, jobs=0, shadow_rays=0,
               radiance_rays=0, launches=0, nodes=0, tris=0, culled=0)
    e0.record()
    t0 = time.perf_counter()
    for i in range(W, W + K):
        st = step(i)
        agg["render_ms"] += st["last_render_ms"]
        agg["shadow_ms"] += st["last_shadow_ms"]
        agg["extend_ms"] += st["last_extend_ms"]
        agg["shadow_launches"] += st["last_shadow_launches"]
        agg["extend_launches"] += st["last_extend_launches"]
        agg["jobs"] += st["last_shadow_jobs"]
        agg["shadow_rays"] += st["last_shadow_rays"]
        agg["radiance_rays"] += st["last_radiance_rays"]
        agg["launches"] += st["last_kernel_launches"]  # our kernels only (k_pool/k_path + k_finalize; wavefront: every stage launch); R.reset() is a memset, the L2 flush is torch's fill
        agg["nodes"] += st["last_nodes_visited"]
        agg["tris"] += st["last_triangles_tested"]
        agg["culled"] += st["last_shadow_culled"]
    torch.cuda.synchronize()
    e1.record()
    e1.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = clk.stop() if clk else None
    # whole device-synchronised bracket; every call inside is blocking, so event time == wall time up to microseconds
    t_ms = torch.tensor([max(e0.elapsed_time(e1), agg["render_ms"])], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    T = float(t_ms.item())
    value = world * npix * S * K / T / 1e3

    # ---- end to end: host buffers in, host image out, every step
    e2e = None
    if not args.no_e2e:
        h2d = sc["vertices"].nbytes + sc["normals"].nbytes + sc["mat_indices"].nbytes + len(sc["materials_packed"])
        d2h = npix * 16
        # pinned host memory for the step's inputs and for the image that comes back
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        sc_pinned = dict(sc, vertices=pin(sc["vertices"]), normals=pin(sc["normals"]), mat_indices=pin(sc["mat_indices"]))
        img_pinned = torch.empty((h, w, 4), dtype=torch.float32).pin_memory().numpy()
        # W untimed warm-up steps here too (at most 3): the first create of a second context allocates its state with
        # cudaMalloc (tens of ms); from then on the library's allocator cache serves it, as in any steady state
        We = min(W, 3)
        t0 = time.perf_counter()
        for i in range(-We, K):
            if i == 0:
                barrier()
                t0 = time.perf_counter()
            ts = [time.perf_counter()]
            R2 = rt.Renderer.from_scene(sc_pinned, device=local_rank, flags=pipe_flag)     # H2D + BVH build
            ts.append(time.perf_counter())
            R2.render_subframes((W + K + i) * world + rank, 1, S)
            ts.append(time.perf_counter())
            if world > 1:
                t2 = ldist.accum_tensor(R2)
                dist.reduce(t2, dst=0, op=dist.ReduceOp.SUM)
                torch.cuda.synchronize()
            if rank == 0:
                img = R2.read_accum(out=img_pinned)                       # D2H
                assert np.isfinite(img[0, 0, 0])
            ts.append(time.perf_counter())
            R2.close()
            ts.append(time.perf_counter())
            if os.environ.get("BENCH_VERBOSE"):
                sys.stderr.write("e2e step %d: create %.1f render %.1f read %.1f destroy %.1f ms\n" % (
                    i, *[(ts[k + 1] - ts[k]) * 1e3 for k in range(4)]))
        barrier()
        te = torch.tensor([(time.perf_counter() - t0) * 1e3], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": round(world * npix * S * K / float(te.item()) / 1e3, 4), "unit": "Msamples/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": round(float(te.item()) / K, 3),
               "warmup": We}

    if rank == 0:
        peak, peak_src = measured_peaks()
        ref_rays = agg["shadow_rays"] + agg["radiance_rays"]          # rays the reference's programs would trace
        rays = ref_rays - agg["culled"]                                # rays actually traversed here
        nn, nt = agg["nodes"] / max(rays, 1), agg["tris"] / max(rays, 1)
        ext_launches = max(agg["extend_launches"], 1)
        avg_ms = agg["extend_ms"] / ext_launches
        if args.pipeline != "wavefront":
            # dominant kernel: k_pool (the default picks it for this workload: 4 M chains per step) or k_path.  Algorithmic bytes (DESIGN.md "Kernels"): per traversed ray 80 B per node visited
            # and 48 B per triangle tested; per radiance hit 48 B of vertex normals + 48 B of material; per chain one
            # 16 B sum written.  All of it but the sums is served by L1/L2 (BVH 11 KB + triangles 96 KB).
            kname, prof_name = ("k_path<wide8>", "r01_k_path.json") if args.pipeline == "path" else ("k_pool<wide8>", "r01_k_pool.json")
            alg_bytes = (rays * (80 * nn + 48 * nt) + agg["radiance_rays"] * 96 + npix * K * 16) / ext_launches
            per_unit = alg_bytes / (npix * S)
            per_unit_name = "algorithmic_bytes_per_sample"
            note = ("one persistent launch per step; chain state never leaves the SM (registers for k_path, shared memory for "
                    "k_pool), BVH (11 KB) + triangles (96 KB) are L1/L2 resident: the kernel is issue bound, HBM sees only the "
                    "16-byte sum per chain. `issue` (issue-slot utilisation x active lanes / 32, committed ncu capture) is the "
                    "roofline that binds.")
        else:
            # k_extend (one radiance ray + material dispatch per live chain and iteration): chain state 16 (sum) + 32 (a, c)
            # + 32 (o, d) read, 64 written, hit shading 96, ray 32, BVH 80 B per node visited and 48 B per triangle tested
            kname, prof_name = "k_extend<wide8>", "r01_k_extend.json"
            chains_per_launch = agg["radiance_rays"] / ext_launches
            per_unit = 144 + 96 + 32 + 80 * nn + 48 * nt
            per_unit_name = "algorithmic_bytes_per_chain"
            alg_bytes = chains_per_launch * per_unit
            note = ("wavefront pipeline (ablation): BVH + triangles are L1/L2 resident, HBM traffic is the chain state; "
                    "latency/issue bound.")
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else None
        traffic, issue = None, None
        prof = os.path.join(ROOT, "profiles", prof_name)
        if os.path.exists(prof):
            try:
                pj = json.load(open(prof))["launches"][0]
                traffic = pj.get("dram_bytes_per_launch")
                issue = {"issue_slot_utilisation_pct": pj.get("issue_slot_utilisation_pct"),
                         "avg_active_lanes_per_instruction": pj.get("avg_active_lanes_per_instruction"),
                         "frac_of_issue_roofline": pj.get("issue_roofline_frac"),
                         "capture": pj.get("capture", "profiles/%s (ncu --set full)" % prof_name)}
            except Exception:
                pass
        out = {
            "metric": METRIC, "value": round(value, 4), "unit": "Msamples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(T / K, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: README Cornell box 2000x2000, 7 bounces, 1002 authored triangles; "
                                   "step = one subframe of %d spp per GPU (2000 spp = %d steps)" % (S, max(1, 2000 // S)),
                       "width": w, "height": h, "bounces": sc["num_bounces"], "spp_per_step": S, "parallelism": "sample-space x%d" % world,
                       "pipeline": args.pipeline, "l2": "flushed between steps (256 MB rewritten; L2 is 126 MB)"},
            "mrays_per_s": round(world * rays / T / 1e3, 1),
            "rays_per_sample": round(rays / (npix * S * K), 2),
            "reference_rays_per_sample": round(ref_rays / (npix * S * K), 2),
            "reference_equivalent_mrays_per_s": round(world * ref_rays / T / 1e3, 1),
            "shadow_tries_resolved_without_traversal": round(agg["culled"] / max(agg["shadow_rays"], 1), 4),
            "wall_ms_per_step": round(wall_ms / K, 3), "device_render_ms_per_step": round(agg["render_ms"] / K, 3),
            "gpu_launches": int(agg["launches"]),
            "clocks": clocks,
            "e2e": e2e,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": round(achieved, 1) if achieved else None, "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 4) if achieved else None, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                         per_unit_name: round(per_unit, 1), "avg_launch_ms": round(avg_ms, 4),
                         "launches": int(ext_launches), "share_of_step": round(agg["extend_ms"] / max(agg["render_ms"], 1e-9), 4),
                         "light_sampling_share_of_step": round(agg["shadow_ms"] / max(agg["render_ms"], 1e-9), 4),
                         "nodes_per_ray": round(nn, 2), "tris_per_ray": round(nt, 2), "issue": issue,
                         "note": note},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(sc, S)
        emit(out)
    R.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
