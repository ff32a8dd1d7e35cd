#!/usr/bin/env python3
"""'Dragon-class' mesh for BASELINE config C3: a closed, smooth, ~870k-triangle tube around a (2,3) torus knot with a
rippled radius, sized to stand inside the Cornell box.  Seedless and deterministic.  Writes an OBJ the reference's
loader accepts ("v", "vn", "f a//a b//b c//c").  Not committed (50 MB): generated on demand into out/.

  python assets/gen_knot.py out/knot.obj [nu nv]      default 1320 x 330 -> 2*nu*nv = 871,200 triangles
"""
import sys

import numpy as np


def knot(nu=1320, nv=330, scale=0.09, centre=(-0.01, -0.02, -0.2), tube=0.22):
    u = np.linspace(0, 2 * np.pi, nu, endpoint=False)
    p, q = 2, 3
    r = np.cos(q * u) + 2.0
    c = np.stack([r * np.cos(p * u), r * np.sin(p * u), -np.sin(q * u)], 1)          # centre line
    d = np.gradient(c, axis=0, edge_order=2)
    d = 0.5 * (np.roll(c, -1, 0) - np.roll(c, 1, 0))
    t = d / np.linalg.norm(d, axis=1, keepdims=True)
    # parallel-transport-free frame: project a fixed axis, good enough for a (2,3) knot (never parallel to z for long)
    a = np.tile(np.array([[0.0, 0.0, 1.0]]), (nu, 1))
    n1 = a - (a * t).sum(1, keepdims=True) * t
    n1 /= np.linalg.norm(n1, axis=1, keepdims=True)
    n2 = np.cross(t, n1)
    v = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    rad = tube * (1.0 + 0.15 * np.sin(8 * u)[:, None] * np.cos(3 * v)[None, :])      # rippled tube radius
    pos = c[:, None, :] + rad[..., None] * (np.cos(v)[None, :, None] * n1[:, None, :] + np.sin(v)[None, :, None] * n2[:, None, :])
    pos = pos * scale / 3.0 + np.array(centre)[None, None, :]
    # smooth normals from the grid (central differences), pointing away from the centre line
    du = np.roll(pos, -1, 0) - np.roll(pos, 1, 0)
    dv = np.roll(pos, -1, 1) - np.roll(pos, 1, 1)
    nrm = np.cross(du, dv)
    nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
    out = pos - (c[:, None, :] * scale / 3.0 + np.array(centre)[None, None, :])
    flip = np.sign((nrm * out).sum(2, keepdims=True))
    nrm *= np.where(flip == 0, 1.0, flip)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a00 = i * nv + j
    a10 = ((i + 1) % nu) * nv + j
    a01 = i * nv + (j + 1) % nv
    a11 = ((i + 1) % nu) * nv + (j + 1) % nv
    faces = np.concatenate([np.stack([a00, a10, a11], -1).reshape(-1, 3), np.stack([a00, a11, a01], -1).reshape(-1, 3)])
    return pos.reshape(-1, 3).astype(np.float32), nrm.reshape(-1, 3).astype(np.float32), faces.astype(np.int64)


def soup_arrays(nu=1320, nv=330):
    """De-indexed soup (3 vertices + 3 normals per triangle) as the C ABI wants it."""
    p, n, f = knot(nu, nv)
    return p[f.reshape(-1)], n[f.reshape(-1)]


def main():
    path = sys.argv[1]
    nu = int(sys.argv[2]) if len(sys.argv) > 2 else 1320
    nv = int(sys.argv[3]) if len(sys.argv) > 3 else 330
    p, n, f = knot(nu, nv)
    with open(path, "w") as fh:
        fh.write("# (2,3) torus-knot tube, %d x %d, %d triangles\n" % (nu, nv, len(f)))
        np.savetxt(fh, p, fmt="v %.9g %.9g %.9g")
        np.savetxt(fh, n, fmt="vn %.9g %.9g %.9g")
        g = f + 1
        np.savetxt(fh, np.stack([g[:, 0], g[:, 0], g[:, 1], g[:, 1], g[:, 2], g[:, 2]], 1), fmt="f %d//%d %d//%d %d//%d")
    print("wrote %s: %d vertices, %d triangles" % (path, len(p), len(f)))


if __name__ == "__main__":
    main()
