"""Python restatement of the reference's scene front-end.  TEST INFRASTRUCTURE.

Follows src/LiSA/src/scene_parser.cc:35-262 (same regular expressions, same
first-match-wins and error rules) and src/LiSA/src/parse_obj.cc:4-69 (single
space splitting, first three face vertices, mandatory normal index).  Pinned
by tests/golden/ref_parse_*.json, which are dumps of the reference's own
parser (oracle/ref_parse_main.cc) on the same files.
"""
import re
import struct

import numpy as np

VAR = r"([a-zA-Z0-9]|_)+"             # scene_parser.hh:43
PATH = r"([a-zA-Z0-9]|_|\/|.)+"       # scene_parser.hh:44  (the '.' is unescaped)
FLOAT = r"\s*([+-]?([0-9]*[.])?[0-9]+)\s*"


class SceneError(Exception):
    """The reference prints the message on stderr and exits 1 (or -1 for a missing OBJ)."""

    def __init__(self, msg, code=1):
        super().__init__(msg)
        self.code = code


def _stof(s):
    return float(np.float32(float(s)))


def _vector_rgx(begin, n):            # scene_parser.cc:66-79
    return re.compile(begin + r"\s*=\s*\(" + ",".join([FLOAT] * n) + r"\)")


def _float_rgx(begin):                # scene_parser.cc:81-83
    return re.compile(begin + r"\s*=\s*([+-]?([0-9]*[.])?[0-9]+)")


def remove_comments(text):            # scene_parser.cc:59-64
    return re.sub(r"\/\*((.|\n)*?)\*\/", "", text)


def parse_obj(path, vertices, normals, mat_indices, mat_idx):   # parse_obj.cc:24-69
    try:
        f = open(path, "r")
    except OSError:
        raise SceneError("%s not found." % path, code=-1)
    vt, nt = [], []
    with f:
        for line in f.read().split("\n"):
            sp = line.split(" ")
            if sp[0] == "v":
                vt.append((_stof(sp[1]), _stof(sp[2]), _stof(sp[3])))
            elif sp[0] == "vn":
                nt.append((_stof(sp[1]), _stof(sp[2]), _stof(sp[3])))
            elif sp[0] == "f":
                for i in range(1, 4):
                    idx = sp[i].split("/")
                    vertices.append(vt[int(idx[0]) - 1])
                    normals.append(nt[int(idx[2]) - 1])
                mat_indices.append(mat_idx)


def pack_material(roughness=0.0, alpha=1.0, n=0.0, diffuse=(0, 0, 0), emit=False, emission=(0, 0, 0)):
    """40-byte Material of src/LiSA/include/structs.hh:16-23 (uninitialised fields zeroed, Q12)."""
    return struct.pack("<3f3fB3x3f", roughness, alpha, n, *diffuse, 1 if emit else 0, *emission)


def parse_scene(path, load_meshes=True):
    try:
        # read_file: getline + "\n" per line (scene_parser.cc:5-19)
        with open(path, "r") as f:
            text = "".join(l + "\n" for l in f.read().splitlines())
    except OSError:
        raise SceneError("%s not found" % path)
    text = remove_comments(text)

    def search_int(name):             # scene_parser.cc:89-103
        m = re.search(name + r"\s*=\s*([0-9]+)", text)
        if not m:
            raise SceneError("Param %s not found." % name)
        return int(m.group(1))

    out = {}
    out["width"] = search_int("width")
    out["height"] = search_int("height")
    out["num_samples"] = search_int("num_samples")
    out["num_bounces"] = search_int("num_bounces")
    m = re.search(r"output_image\s*=\s*(" + PATH + ")", text)
    if not m:
        raise SceneError("Param output_image not found.")
    out["output_image"] = m.group(1)

    # materials (scene_parser.cc:106-163)
    blocks = [m.group(0) for m in re.finditer(r"material\s+" + VAR + r"+\s*\{\n*[^\}]*", text)]
    if not blocks:
        raise SceneError("Error in materials declaration: no valid materials declared.")
    names, mats = {}, []
    for s in blocks:
        m = re.search(r"\s+(" + VAR + r")\s*\{", s)
        name = m.group(1) if m else ""
        is_light = re.search(r"(emit\s*=\s*true)", s) is not None
        m = _vector_rgx("color", 4).search(s)
        if not m:
            raise SceneError("Error in materials declaration: no valid color provided in material %s." % name)
        col = [_stof(m.group(1)), _stof(m.group(3)), _stof(m.group(5))]
        alpha = _stof(m.group(7))
        if col[0] > 1 or col[1] > 1 or col[2] > 1:
            col = [float(np.float32(np.float64(c) / 255.0)) for c in col]
        if is_light:
            mat = dict(emit=True, alpha=1.0, emission=tuple(col))
        else:
            m = _float_rgx("(roughness | n)").search(s)
            if not m:
                raise SceneError("Error in materials declaration: no valid roughness|refractive index in material %s." % name)
            p = _stof(m.group(2))
            mat = dict(emit=False, alpha=alpha, diffuse=tuple(col))
            if alpha < 1.0:
                mat["n"] = p
            else:
                mat["roughness"] = p
        if name not in names:
            mats.append(mat)
            names[name] = len(mats) - 1
    out["materials"] = mats
    out["material_names"] = names

    # meshes (scene_parser.cc:165-198)
    verts, norms, midx, files = [], [], [], []
    for m0 in re.finditer(r"mesh\s*\{\n*[^\}]*", text):
        s = m0.group(0)
        m = re.search(r"material\s*=\s*(" + VAR + ")", s)
        if not m:
            raise SceneError("Error in mesh declaration: no valid material provided")
        if m.group(1) not in names:
            raise SceneError("Error in mesh declaration: %s material not found" % m.group(1))
        mi = names[m.group(1)]
        m = re.search(r"obj_file\s*=\s*(.+\.obj)", s)
        if not m:
            raise SceneError("Error in mesh declaration: no valid obj_file provided")
        files.append((m.group(1), mi))
        if load_meshes:
            parse_obj(m.group(1), verts, norms, midx, mi)
    out["mesh_files"] = files

    # camera (scene_parser.cc:200-245)
    cams = [m.group(0) for m in re.finditer(r"camera\s*\{\n*[^\}]*", text)]
    if len(cams) != 1:
        raise SceneError("Error in camera declaration: no valid camera declared.")
    s = cams[0]
    m = _vector_rgx("position", 3).search(s)
    if not m:
        raise SceneError("Error in camera declaration: no valid position provided.")
    eye = (_stof(m.group(1)), _stof(m.group(3)), _stof(m.group(5)))
    m = _vector_rgx("look_at", 3).search(s)
    if not m:
        raise SceneError("Error in camera declaration: no valid look_at provided.")
    look = (_stof(m.group(1)), _stof(m.group(3)), _stof(m.group(5)))
    m = _float_rgx("fov").search(s)
    if not m:
        raise SceneError("Error in camera declaration: no valid fov provided.")
    out["camera"] = dict(eye=eye, look_at=look, fov=_stof(m.group(1)))

    out["vertices"] = np.asarray(verts, dtype=np.float32).reshape(-1, 3)
    out["normals"] = np.asarray(norms, dtype=np.float32).reshape(-1, 3)
    out["mat_indices"] = np.asarray(midx, dtype=np.int32)
    out["materials_packed"] = b"".join(
        pack_material(roughness=mm.get("roughness", 0.0), alpha=mm["alpha"], n=mm.get("n", 0.0),
                      diffuse=mm.get("diffuse", (0, 0, 0)), emit=mm["emit"], emission=mm.get("emission", (0, 0, 0)))
        for mm in mats)
    return out
