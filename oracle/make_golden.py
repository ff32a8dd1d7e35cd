#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference itself.  TEST INFRASTRUCTURE; run in the build container.

  ref_kat.json         output of oracle/_ref/lisa_ref_kat  (reference headers compiled on the host)
  ref_parse_*.json     output of oracle/_ref/lisa_ref_parse (reference SceneParser + parse_obj) on
                       scenes/*.rto and tests/golden/parser_cases/*.rto
  optix_*.npz          accumulators dumped by the UNMODIFIED reference OptiX renderer
                       (oracle/_ref/lisa_optix_ref) on a B200; the raw dumps are produced on the GPU box:
                         gpurun -- 'R=oracle/_ref/lisa_optix_ref; mkdir -p out;
                           $R -s scenes/cornell_tiny.rto --spp 1    --accum gpurun_out/optix_tiny_1.f32;
                           $R -s scenes/cornell_tiny.rto            --accum gpurun_out/optix_tiny_16.f32;
                           $R -s scenes/cornell_tiny.rto --spp 1024 --accum gpurun_out/optix_tiny_1024.f32;
                           $R -s scenes/cornell_tiny.rto --spp 4 --subframes 4 --accum gpurun_out/optix_tiny_4x4.f32;
                           $R -s scenes/cornell_c1.rto --spp 1      --accum gpurun_out/optix_c1_1.f32;
                           $R -s scenes/cornell_c1.rto              --accum gpurun_out/optix_c1.f32;
                           $R -s scenes/cornell_c1.rto --spp 1024   --accum gpurun_out/optix_c1_1024.f32'
                       and this script reduces them to small fixtures (full image for 64x64, an
                       8x8-block mean and a 128x128 centre crop for 512x512).
"""
import glob
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    os.makedirs(G, exist_ok=True)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    kat = subprocess.check_output([os.path.join(REF, "lisa_ref_kat")])
    json.loads(kat)
    open(os.path.join(G, "ref_kat.json"), "wb").write(kat)
    scenes = sorted(glob.glob(os.path.join(ROOT, "scenes", "*.rto")) + glob.glob(os.path.join(G, "parser_cases", "*.rto")))
    for s in scenes:
        name = os.path.splitext(os.path.basename(s))[0]
        out = os.path.join(G, "ref_parse_%s.json" % name)
        r = subprocess.run([os.path.join(REF, "lisa_ref_parse"), os.path.relpath(s, ROOT), out], cwd=ROOT,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if r.returncode != 0:  # the reference rejects the file: record how
            json.dump({"error": r.stderr.decode().strip(), "exit": r.returncode}, open(out, "w"))
        else:
            json.load(open(out))
    dumps = {"optix_tiny_1": (64, 64), "optix_tiny_16": (64, 64), "optix_tiny_1024": (64, 64), "optix_tiny_4x4": (64, 64),
             "optix_c1_1": (512, 512), "optix_c1": (512, 512), "optix_c1_1024": (512, 512), "optix_c1_16x4": (512, 512)}
    for name, (w, h) in dumps.items():
        p = os.path.join(ROOT, "gpurun_out", name + ".f32")
        if not os.path.exists(p):
            continue
        a = np.fromfile(p, dtype=np.float32).reshape(h, w, 4)[..., :3]
        out = {"mean_rgb": a.reshape(-1, 3).mean(axis=0, dtype=np.float64)}
        if w <= 64:
            out["accum"] = a
        else:
            out["block8"] = a.reshape(h // 8, 8, w // 8, 8, 3).mean(axis=(1, 3), dtype=np.float64).astype(np.float32)
            out["crop"] = a[h // 2 - 64:h // 2 + 64, w // 2 - 64:w // 2 + 64].copy()
            out["crop_origin"] = np.array([h // 2 - 64, w // 2 - 64])
        np.savez_compressed(os.path.join(G, name + ".npz"), **out)
        print(name, out["mean_rgb"])


if __name__ == "__main__":
    sys.exit(main())
