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
  roofline  dominant kernel k_pool (the whole estimator, one persistent launch per step).  What binds it is instruction
          ISSUE (L2-resident scene), so: achieved = thread instructions per launch / the launch's
          duration measured live (CUDA events around the launch, LISA_FLAG_PROFILE_STAGES); peak = SMs x 4 schedulers
          x 32 lanes x the SM clock sampled during the timed region; frac = issue-slot utilisation x active lanes / 32.
          The instruction counts per launch are a property of (kernel, workload): they come from the ncu capture of
          the same launch committed under profiles/, which records a hash of the kernel sources — `stale` says whether
          the sources have changed since.  HBM traffic of the launch is reported beside it (secondary).
  multi_gpu_parity  (N > 1) after the timed loop a 128x128 job is rendered partitioned over the N ranks + reduced, and
          again by rank 0 alone over the same subframes; the line carries the largest difference.
  strong_scaling    the same workload with a FIXED total of samples per step split over the N ranks.
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
import hashlib
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.chdir(ROOT)

SCENE = "scenes/cornell_c2.rto"
METRIC = "Msamples/s (Cornell 2000x2000, 7 bounces)"


STRONG_SPP = 48          # strong-scaling leg: samples per pixel and step over ALL ranks (divisible by 1, 2, 4, 8)


def scene_header(path=SCENE):
    """width / height / num_bounces straight from the scene text (the grammar's `name = <uint>`, scene_parser.cc:85-87):
    all the reference arm needs to name its config, without loading any library."""
    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, path)).read(), flags=re.S)
    g = lambda name: int(re.search(name + r"\s*=\s*([0-9]+)", txt).group(1))
    return g("width"), g("height"), g("num_bounces")


def config_dict(w, h, bounces, S):
    """The workload, byte-identical in both arms."""
    return {"workload": "BASELINE configs[1]: README Cornell box %dx%d, %d bounces, 1002 authored triangles; step = one subframe "
                        "of %d spp per GPU (2000 spp = %d steps)" % (w, h, bounces, S, max(1, 2000 // S)),
            "width": w, "height": h, "bounces": bounces, "spp_per_step": S,
            "l2": "flushed between steps (256 MB rewritten; L2 is 126 MB)"}


def kernel_source_hash():
    """sha256 over the sources the render kernels are built from (what an ncu capture under profiles/ is valid for)."""
    h = hashlib.sha256()
    base = os.path.join(ROOT, "lisa_b200")
    names = ["Makefile"] + sorted(os.path.join("csrc", f) for f in os.listdir(os.path.join(base, "csrc")) if f.endswith((".cuh", ".h")) or f == "estimator.cu")
    names += [os.path.join("csrc", "bsdf", "lambertian.cuh")]
    for n in names:
        h.update(n.encode())
        h.update(open(os.path.join(base, n), "rb").read())
    return h.hexdigest()[:16]


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
    """Reference arm: the unmodified reference's OptiX renderer (it has no CPU implementation).  Nothing of the product is
    imported or loaded here: the config comes from the scene text, the numbers from the harness's own JSON line."""
    if rank != 0:
        return 0
    w, h, bounces = scene_header()
    S, K, W = args.spp_per_step, args.steps, args.warmup
    exe = os.path.join(ROOT, "oracle", "_ref", "lisa_optix_ref")
    base = {"metric": METRIC, "unit": "Msamples/s", "n_gpus": 1, "steps": K, "warmup": W, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": config_dict(w, h, bounces, S)}
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
                assert (j["width"], j["height"], j["bounces"], j["spp"]) == (w, h, bounces, S), j
                ms = sum(j["step_ms"])
                v = w * h * S * K / ms / 1e3
                out = dict(base, value=round(v, 4), ms_per_step=round(ms / K, 3), clocks=clocks, gpu_launches=K,
                           cpu_baseline={"value": round(v, 4), "unit": "Msamples/s", "cores": 1, "kind": "reference",
                                         "sample": "unmodified gaetanserre/LiSA OptiX 7.4 renderer (oracle/_ref/lisa_optix_ref) on "
                                                   "1 B200 (software traversal, no RT cores), 1 host thread; %d launches of %d spp, "
                                                   "std::chrono around optixLaunch+sync; setup (upload + optixAccelBuild + pipeline) %.1f ms"
                                                   % (K, S, j.get("setup_ms", 0.0))},
                           e2e={"value": round(v, 4), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
            else:
                sys.stderr.write("reference OptiX harness failed (rc %d): %s\n" % (r.returncode, r.stderr[-400:]))
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("reference OptiX harness failed: %r\n" % (e,))
    if out is None:
        # OptiX could not start: time the CPU restatement instead (labelled as a port); scene through the oracle's own parser
        from oracle import scene_py
        sys.stderr.write("reference arm: OptiX unavailable, timing the CPU restatement (oracle/) on %s\n" % SCENE)
        sc = scene_py.parse_scene(SCENE)
        cb = cpu_baseline(sc, S, target_s=float(os.environ.get("LISA_BENCH_CPU_TARGET_S", "20")))  # env: the CPU tests shorten it
        out = dict(base, value=cb["value"], ms_per_step=round(w * h * S / (cb["value"] * 1e3), 1), cpu_baseline=cb, gpu_launches=0,
                   clocks={"sm_mhz": None, "sm_max_mhz": None, "reasons": []},
                   e2e={"value": cb["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    # evidence for the driver's hygiene check: this arm must not map any of the product's libraries
    mapped = sorted({l.split()[-1] for l in open("/proc/self/maps") if "lisa_b200" in l}) if os.path.exists("/proc/self/maps") else []
    sys.stderr.write("reference arm: product libraries mapped in this process: %s; product modules imported: %s\n"
                     % (mapped, sorted(m for m in sys.modules if m.startswith("lisa_b200"))))
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="which leg is the line's `value`: weak = one subframe of --spp-per-step per GPU and step (default), "
                         "strong = a fixed %d spp per step split over the GPUs; the other leg is reported beside it" % STRONG_SPP)
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
    assert (w, h, sc["num_bounces"]) == scene_header()
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

    def step(i, spp):
        """One pass of the hot path: this rank's subframe of step i (+ the one reduce when N > 1)."""
        R.reset()
        flush.fill_(i & 0xff)
        torch.cuda.current_stream().synchronize()
        R.render_subframes(i * world + rank, 1, spp)
        if world > 1:
            dist.reduce(acc_t, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                total.add_(acc_t)
        return R.stats()

    def timed_leg(first_step, spp, sample_clocks):
        """K timed steps of `spp` samples per pixel and rank, bracketed by barrier + synchronize; device time, max over ranks."""
        barrier()
        clk = ClockSampler(local_rank) if sample_clocks else None   # every rank watches its own GPU
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        agg = dict(render_ms=0.0, shadow_ms=0.0, extend_ms=0.0, shadow_launches=0, extend_launches=0, jobs=0, shadow_rays=0,
                   radiance_rays=0, launches=0, nodes=0, tris=0, culled=0)
        e0.record()
        t0 = time.perf_counter()
        for i in range(first_step, first_step + K):
            st = step(i, spp)
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
        per_rank = None
        if world > 1 and sample_clocks:
            # what each rank saw: the slowest one sets the line's time, and its clocks / power say why
            mine = {"rank": rank, "ms_per_step": round(float(e0.elapsed_time(e1)) / K, 3), "render_ms_per_step": round(agg["render_ms"] / K, 3),
                    "sm_mhz": (clocks or {}).get("sm_mhz"), "power_w_max": (clocks or {}).get("power_w_max"), "reasons": (clocks or {}).get("reasons")}
            per_rank = [None] * world
            dist.all_gather_object(per_rank, mine)
        return dict(T=T, value=world * npix * spp * K / T / 1e3, wall_ms=wall_ms, agg=agg, clocks=clocks, per_rank=per_rank)

    # ---- weak leg: every rank renders its own subframe of S spp each step
    for i in range(W):
        step(i, S)
    weak = timed_leg(W, S, True)
    # ---- strong leg: a FIXED total of STRONG_SPP samples per pixel and step, split over the ranks (each renders its own
    # subframe of STRONG_SPP / N spp).  At N = 1 it is the weak leg with another spp.
    strong_spp = max(1, STRONG_SPP // world)
    for i in range(min(W, 2)):
        step(W + K + i, strong_spp)
    strong = timed_leg(W + K + 2, strong_spp, False)
    strong_obj = {"value": round(strong["value"], 4), "unit": "Msamples/s", "ms_per_step": round(strong["T"] / K, 3),
                  "spp_per_step_all_gpus": strong_spp * world, "spp_per_step_per_gpu": strong_spp,
                  "note": "fixed total work per step split over the ranks (sample-space), one reduce per step"}
    weak_obj = {"value": round(weak["value"], 4), "unit": "Msamples/s", "ms_per_step": round(weak["T"] / K, 3), "spp_per_step_per_gpu": S}
    main = strong if args.scaling == "strong" else weak
    T, value, agg, wall_ms, clocks = main["T"], main["value"], main["agg"], main["wall_ms"], weak["clocks"]
    S_main = strong_spp if args.scaling == "strong" else S

    # ---- multi-GPU parity, where the driver can see it (its GPU-test box has one GPU): the N-rank partition + reduce of a
    # small job against rank 0 rendering the same subframes alone
    parity = None
    if world > 1:
        small = dict(sc, width=128, height=128)
        first, count, spp_p = 5, 2 * world + 1, 4      # ragged: the blocks differ in size
        Rp = rt.Renderer.from_scene(small, device=local_rank, flags=pipe_flag)
        ldist.render_partitioned(Rp, first, count, spp_p)
        if rank == 0:
            multi = Rp.read_accum()
            Rs = rt.Renderer.from_scene(small, device=local_rank, flags=pipe_flag)
            Rs.render_subframes(first, count, spp_p)
            single = Rs.read_accum()
            Rs.close()
            diff = float(np.abs(multi - single).max())
            # same samples, another order of the float additions: 2e-6 relative (tests/test_gpu_multi.py uses the same bar)
            ok = bool(np.allclose(multi, single, rtol=2e-6, atol=1e-7)) and bool((single[..., :3] > 0).mean() > 0.5)
            parity = {"max_abs_diff": diff, "ok": ok, "ranks": world, "job": "128x128, subframes [%d, %d) x %d spp in contiguous blocks per rank, one NCCL reduce" % (first, first + count, spp_p),
                      "image_mean": float(single[..., :3].mean())}
        Rp.close()
        barrier()

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
            R2.render_subframes((W + 3 * K + 8 + i) * world + rank, 1, S_main)
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
        e2e = {"value": round(world * npix * S_main * K / float(te.item()) / 1e3, 4), "unit": "Msamples/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": round(float(te.item()) / K, 3),
               "warmup": We}

    if rank == 0:
        hbm_peak, hbm_src = measured_peaks()
        ref_rays = agg["shadow_rays"] + agg["radiance_rays"]          # rays the reference's programs would trace
        rays = ref_rays - agg["culled"]                                # rays actually traversed here
        nn, nt = agg["nodes"] / max(rays, 1), agg["tris"] / max(rays, 1)
        ext_launches = max(agg["extend_launches"], 1)
        avg_ms = agg["extend_ms"] / ext_launches
        if args.pipeline != "wavefront":
            # dominant kernel: k_pool (the default picks it for this workload: 4 M chains per step) or k_path
            kname, prof_name = ("k_path<wide8>", "r01_k_path.json") if args.pipeline == "path" else ("k_pool<wide8>", "r02_k_pool.json")
            # L1-level algorithmic bytes (DESIGN.md "Kernels"): per traversed ray 80 B per node visited and 48 B per triangle
            # tested; per radiance hit 48 B of vertex normals + 48 B of material; per chain one 16 B sum written
            alg_bytes = (rays * (80 * nn + 48 * nt) + agg["radiance_rays"] * 96 + npix * K * 16) / ext_launches
        else:
            kname, prof_name = "k_extend<wide8>", "r01_k_extend.json"
            alg_bytes = (agg["radiance_rays"] / ext_launches) * (144 + 96 + 32 + 80 * nn + 48 * nt)
        # ---- the roofline that binds: instruction issue.  Instruction counts per launch from the committed ncu capture of
        # this very launch (one step of S spp), valid while the kernel sources hash the same; its duration and the SM clock are
        # measured in THIS run.
        # the newest capture of that kernel under profiles/ (r03_ = second session of round 2, r02_, r01_)
        cands = [os.path.join(ROOT, "profiles", tag + prof_name[4:]) for tag in ("r03_", "r02_", "r01_")]
        prof, pj = next((c for c in cands if os.path.exists(c)), cands[-1]), None
        if os.path.exists(prof):
            try:
                pj = json.load(open(prof))["launches"][0]
            except Exception:  # noqa: BLE001
                pj = None
        sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
        sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        peak_thread = sm_count * 4 * 32 * sm_mhz * 1e6                 # thread instructions / s
        src_hash = kernel_source_hash()
        roof = {"bound": "issue", "kernel": kname, "achieved": None, "peak": round(peak_thread / 1e12, 4), "unit": "Tthread-inst/s",
                "frac": None, "traffic": None,
                "peak_source": "%d SMs x 4 schedulers x 32 lanes x %.0f MHz (SM clock sampled by nvidia-smi during the timed region)" % (sm_count, sm_mhz),
                "avg_launch_ms": round(avg_ms, 4), "launches": int(ext_launches),
                "share_of_step": round(agg["extend_ms"] / max(agg["render_ms"], 1e-9), 4),
                "nodes_per_ray": round(nn, 2), "tris_per_ray": round(nt, 2)}
        if pj and pj.get("warp_instructions") and avg_ms > 0 and args.scaling == "weak" and S == int(pj.get("spp_per_launch", 50)):
            warp_i = float(pj["warp_instructions"])
            thread_i = float(pj.get("thread_instructions") or warp_i * pj["avg_active_lanes_per_instruction"])
            ach = thread_i / (avg_ms * 1e-3)
            roof.update(achieved=round(ach / 1e12, 4), frac=round(ach / peak_thread, 4), traffic=pj.get("dram_bytes_per_launch"),
                        issue={"warp_instructions_per_launch": warp_i, "thread_instructions_per_launch": thread_i,
                               "issue_slot_utilisation": round(warp_i / (avg_ms * 1e-3) / (sm_count * 4 * sm_mhz * 1e6), 4),
                               "active_lanes_per_instruction": round(thread_i / warp_i, 2),
                               "source": "profiles/%s (ncu --set full of one launch of %d spp)" % (os.path.basename(prof), S),
                               "source_kernel_hash": pj.get("kernel_source_sha256"), "kernel_hash": src_hash,
                               "stale": pj.get("kernel_source_sha256") != src_hash,
                               "under_ncu": {"duration_ms": pj.get("duration"), "issue_slot_utilisation_pct": pj.get("issue_slot_utilisation_pct"),
                                             "frac": pj.get("issue_roofline_frac")}})
            if roof["traffic"]:
                gbs = roof["traffic"] / (avg_ms * 1e-3) / 1e9
                roof["hbm"] = {"achieved": round(gbs, 2), "peak": hbm_peak, "unit": "GB/s", "frac": round(gbs / hbm_peak, 6), "peak_source": hbm_src,
                               "note": "HBM sees the 16-byte sum per chain and the first touch of the scene: not the bound"}
        roof["l1_level_algorithmic_bytes_per_launch"] = int(alg_bytes)
        roof["note"] = ("one persistent launch per step; chain state never leaves the SM, BVH (11 KB) + triangles (96 KB) are L1/L2 "
                        "resident: instruction issue binds (SURVEY.md 8d), frac = issue-slot utilisation x active lanes / 32")
        cfg = config_dict(w, h, sc["num_bounces"], S)
        out = {
            "metric": METRIC, "value": round(value, 4), "unit": "Msamples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(T / K, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "parallelism": "sample-space x%d" % world, "pipeline": args.pipeline,
            "weak_scaling": weak_obj, "strong_scaling": strong_obj,
            "mrays_per_s": round(world * rays / T / 1e3, 1),
            "rays_per_sample": round(rays / (npix * S_main * K), 2),
            "reference_rays_per_sample": round(ref_rays / (npix * S_main * K), 2),
            "reference_equivalent_mrays_per_s": round(world * ref_rays / T / 1e3, 1),
            "shadow_tries_resolved_without_traversal": round(agg["culled"] / max(agg["shadow_rays"], 1), 4),
            "wall_ms_per_step": round(wall_ms / K, 3), "device_render_ms_per_step": round(agg["render_ms"] / K, 3),
            "gpu_launches": int(agg["launches"]),
            "clocks": clocks,
            **({"per_rank": weak["per_rank"]} if weak.get("per_rank") else {}),
            "e2e": e2e,
            "roofline": roof,
        }
        if parity is not None:
            out["multi_gpu_parity"] = parity
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
