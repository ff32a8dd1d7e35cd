#!/usr/bin/env python3
"""Author the Cornell-box OBJ meshes the reference's README scene refers to.

The reference repository ships no assets (README.md:74-116 names
assets/objs/cornell_box/{bot,top,back,right,left,large_box,small_box,sphere,light}.obj
but none exist), so the geometry is authored here once and used unchanged by
the oracle, the OptiX reference harness and the CUDA path.

Constraints honoured:
  * camera of the README scene: eye (-0.01, 0.015, 0.6) -> (-0.01, 0.015, 0),
    vertical fov 40 deg => visible half extent at z=0 is 0.6*tan(20deg)=0.2184,
    so the box has half size 0.2, is centred on (-0.01, 0.015, -0.2) and is
    open towards +z;
  * OBJ subset of the reference loader (src/LiSA/src/parse_obj.cc:24-69):
    single-space separated "v", "vn", "f a//c a//c a//c", triangles only,
    normal index mandatory, 1-based indices local to the file;
  * shading normals are never face-forwarded (shader.cu:223), so walls carry
    inward normals and closed objects outward normals.

Run:  python assets/gen_cornell.py   (writes assets/objs/cornell_box/*.obj)
"""
import math
import os

CX, CY, CZ, H = -0.01, 0.015, -0.2, 0.2
X0, X1 = CX - H, CX + H
Y0, Y1 = CY - H, CY + H
Z0, Z1 = CZ - H, CZ + H
# objects float 1e-3 above the floor: coplanar faces would make the closest hit a coin flip between
# implementations (z-fighting), and 1e-3 is 10x the reference's tmin (shader.cu:114)
LIFT = 1e-3


def fmt(x):
    return "%.9g" % (x + 0.0)


def write_obj(path, verts, norms, faces, comment):
    """faces: list of ((v0, n0), (v1, n1), (v2, n2)) with 0-based indices."""
    with open(path, "w") as f:
        f.write("# %s\n" % comment)
        for v in verts:
            f.write("v %s %s %s\n" % tuple(fmt(c) for c in v))
        for n in norms:
            f.write("vn %s %s %s\n" % tuple(fmt(c) for c in n))
        for tri in faces:
            f.write("f " + " ".join("%d//%d" % (v + 1, n + 1) for v, n in tri) + "\n")


def quad(p0, p1, p2, p3, n):
    """Two triangles (p0,p1,p2), (p0,p2,p3) sharing one normal n."""
    verts = [p0, p1, p2, p3]
    faces = [((0, 0), (1, 0), (2, 0)), ((0, 0), (2, 0), (3, 0))]
    return verts, [n], faces


def box(cx, cz, w, h, d, y0, angle_deg):
    """Closed box standing on y0, rotated about +y, outward face normals."""
    a = math.radians(angle_deg)
    ca, sa = math.cos(a), math.sin(a)

    def rot(x, z):
        return (cx + ca * x + sa * z, cz - sa * x + ca * z)

    hw, hd = w / 2, d / 2
    corners = [(-hw, -hd), (hw, -hd), (hw, hd), (-hw, hd)]
    verts = []
    for y in (y0, y0 + h):
        for (x, z) in corners:
            rx, rz = rot(x, z)
            verts.append((rx, y, rz))
    # local face normals, rotated
    def rn(x, z):
        return (ca * x + sa * z, 0.0, -sa * x + ca * z)

    norms = [(0, -1, 0), (0, 1, 0), rn(0, -1), rn(1, 0), rn(0, 1), rn(-1, 0)]
    # vertex ids: bottom 0..3, top 4..7 (same corner order)
    quads = [
        ((0, 1, 2, 3), 0),  # bottom (seen from below)
        ((4, 7, 6, 5), 1),  # top
        ((0, 4, 5, 1), 2),  # -z side
        ((1, 5, 6, 2), 3),  # +x side
        ((2, 6, 7, 3), 4),  # +z side
        ((3, 7, 4, 0), 5),  # -x side
    ]
    faces = []
    for (a0, a1, a2, a3), n in quads:
        faces.append(((a0, n), (a1, n), (a2, n)))
        faces.append(((a0, n), (a2, n), (a3, n)))
    return verts, norms, faces


def uv_sphere(c, r, nu=32, nv=16):
    """UV sphere, smooth outward normals, nu*(2 + 2*(nv-2)) = 960 triangles."""
    verts, norms = [], []

    def add(theta, phi):
        n = (math.sin(theta) * math.cos(phi), math.cos(theta), math.sin(theta) * math.sin(phi))
        norms.append(n)
        verts.append((c[0] + r * n[0], c[1] + r * n[1], c[2] + r * n[2]))
        return len(verts) - 1

    top = add(0.0, 0.0)
    rings = []
    for j in range(1, nv):
        theta = math.pi * j / nv
        rings.append([add(theta, 2 * math.pi * i / nu) for i in range(nu)])
    bot = add(math.pi, 0.0)
    faces = []

    def tri(a, b, c_):
        faces.append(((a, a), (b, b), (c_, c_)))

    for i in range(nu):
        i1 = (i + 1) % nu
        tri(top, rings[0][i1], rings[0][i])
        for j in range(nv - 2):
            a, b = rings[j][i], rings[j][i1]
            c_, d = rings[j + 1][i], rings[j + 1][i1]
            tri(a, b, d)
            tri(a, d, c_)
        tri(bot, rings[-1][i], rings[-1][i1])
    return verts, norms, faces


def main():
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "objs", "cornell_box")
    os.makedirs(out, exist_ok=True)
    W = lambda name, geo, what: write_obj(os.path.join(out, name), *geo, comment=what)

    W("bot.obj", quad((X0, Y0, Z1), (X1, Y0, Z1), (X1, Y0, Z0), (X0, Y0, Z0), (0, 1, 0)), "floor, normal +y")
    # The ceiling is a frame around a hole of the light's footprint.  The reference asks OptiX for the first-FOUND hit
    # of a shadow ray (shader.cu:69); with a ceiling 1 mm behind the light that is, depending on the BVH OptiX
    # happens to build, sometimes the ceiling, and the light sample is lost (measured: 4 % of the image mean on the
    # 871k-triangle scene, none on this one).  With nothing behind the emitter the reference's image is well defined.
    lh = 0.05
    strips = [((X0, Z0), (CX - lh, Z1)), ((CX + lh, Z0), (X1, Z1)), ((CX - lh, Z0), (CX + lh, CZ - lh)), ((CX - lh, CZ + lh), (CX + lh, Z1))]
    tv, tf = [], []
    for (xa, za), (xb, zb) in strips:
        b = len(tv)
        tv += [(xa, Y1, za), (xb, Y1, za), (xb, Y1, zb), (xa, Y1, zb)]
        tf += [((b, 0), (b + 1, 0), (b + 2, 0)), ((b, 0), (b + 2, 0), (b + 3, 0))]
    W("top.obj", (tv, [(0, -1, 0)], tf), "ceiling (frame around the light's footprint), normal -y")
    # The ORIGINAL ceiling (one quad, 1 mm behind the light), kept as a fixture: scenes/c3_knot_q2.rto uses it to pin the
    # measured sensitivity of Q2 (first-found vs closest shadow hits) against the reference's OptiX build.
    W("top_full.obj", quad((X0, Y1, Z0), (X1, Y1, Z0), (X1, Y1, Z1), (X0, Y1, Z1), (0, -1, 0)), "ceiling as first authored (full quad behind the light), normal -y")
    W("back.obj", quad((X0, Y0, Z0), (X1, Y0, Z0), (X1, Y1, Z0), (X0, Y1, Z0), (0, 0, 1)), "back wall, normal +z")
    W("right.obj", quad((X1, Y0, Z0), (X1, Y0, Z1), (X1, Y1, Z1), (X1, Y1, Z0), (-1, 0, 0)), "right wall, normal -x")
    W("left.obj", quad((X0, Y0, Z1), (X0, Y0, Z0), (X0, Y1, Z0), (X0, Y1, Z1), (1, 0, 0)), "left wall, normal +x")
    W("large_box.obj", box(-0.08, -0.27, 0.12, 0.24, 0.12, Y0 + LIFT, 17.0), "tall block")
    W("small_box.obj", box(0.07, -0.13, 0.12, 0.12, 0.12, Y0 + LIFT, -17.0), "short block (glass in the README scene)")
    W("sphere.obj", uv_sphere((-0.10, Y0 + 0.05 + LIFT, -0.09), 0.05), "uv sphere 32x16, smooth normals")
    ly = Y1 - 1e-3
    W("light.obj", quad((CX - lh, ly, CZ - lh), (CX + lh, ly, CZ - lh), (CX + lh, ly, CZ + lh), (CX - lh, ly, CZ + lh),
                        (0, -1, 0)), "area light 0.1 x 0.1 just below the ceiling, normal -y")


if __name__ == "__main__":
    main()
