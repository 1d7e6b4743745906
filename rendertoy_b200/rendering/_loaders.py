"""Wavefront OBJ loader (reference: rendering/_loaders.py, which delegates parsing to pywavefront).

pywavefront is not available on the B200 image, so the parser lives here and reproduces what load_obj
consumed from it (_loaders.py:8-33): per mesh, the FIRST material's interleaved, face-corner-expanded vertex
list (triangle soup in file order; polygons fan-triangulated as (v0, v[i-1], v[i])), with vertex_format one of
V3F / N3F_V3F / T2F_V3F / T2F_N3F_V3F.  pywavefront's version is unpinned by the reference, so this behaviour is
anchored on that call site only ("parity unpinned").

load_obj parses with the native one-pass parser (rt_obj_load in csrc/rt_obj.cu: host code over the mapped file, ~25x the
line-by-line Python formulation below, which a dragon-class file keeps busy for seconds).  _parse_obj / _load_obj_python stay
as the readable definition of the rules; tests compare the two on files that exercise every rule.

After the scatter into MeshVertex rows the positions get the reference's scalar min/max normalisation
(_loaders.py:34-38) in float32:  P = (P - min) / (max - min) - 0.5.
"""
import os

import numpy as np

from ._core import create_buffer, mapped
from ._modeling import Mesh, MeshVertex


class _Material:
    def __init__(self, name):
        self.name = name
        self.vertex_format = None
        self.corners = []  # (vi, ti, ni) per emitted corner, 0-based, -1 when absent


class _ObjMesh:
    def __init__(self, name):
        self.name = name
        self.materials = []
        self.n_faces = 0


def _parse_obj(path):
    pos, nrm, tex = [], [], []
    meshes, materials = [], {}
    mesh = material = None

    def resolve(tok, count):
        i = int(tok)
        return i - 1 if i > 0 else count + i

    with open(path, "r", errors="replace") as fh:
        for line in fh:
            if not line or line[0] in "#\n\r":
                continue
            parts = line.split()
            if not parts:
                continue
            key = parts[0]
            if key == "v":
                pos.append((float(parts[1]), float(parts[2]), float(parts[3])))
            elif key == "vn":
                nrm.append((float(parts[1]), float(parts[2]), float(parts[3])))
            elif key == "vt":
                tex.append((float(parts[1]), float(parts[2]) if len(parts) > 2 else 0.0))
            elif key == "o":
                mesh = _ObjMesh(parts[1] if len(parts) > 1 else None)
                meshes.append(mesh)
            elif key == "usemtl":
                name = parts[1] if len(parts) > 1 else None
                material = materials.setdefault(name, _Material(name))
                if mesh is not None and material not in mesh.materials:
                    mesh.materials.append(material)
            elif key == "f":
                if mesh is None:
                    mesh = _ObjMesh(None)
                    meshes.append(mesh)
                if material is None:
                    material = materials.setdefault("default0", _Material("default0"))
                if material not in mesh.materials:
                    mesh.materials.append(material)
                corners = []
                for tok in parts[1:]:
                    f = tok.split("/")
                    vi = resolve(f[0], len(pos))
                    ti = resolve(f[1], len(tex)) if len(f) > 1 and f[1] else -1
                    ni = resolve(f[2], len(nrm)) if len(f) > 2 and f[2] else -1
                    corners.append((vi, ti, ni))
                if material.vertex_format is None:
                    has_t, has_n = corners[0][1] >= 0, corners[0][2] >= 0
                    material.vertex_format = "_".join((["T2F"] if has_t else []) + (["N3F"] if has_n else []) + ["V3F"])
                for i in range(2, len(corners)):      # fan: (c0, c[i-1], c[i])
                    material.corners += (corners[0], corners[i - 1], corners[i])
                    mesh.n_faces += 1
    return (np.asarray(pos, np.float32).reshape(-1, 3), np.asarray(nrm, np.float32).reshape(-1, 3),
            np.asarray(tex, np.float32).reshape(-1, 2), meshes)


def _normalise_positions(rows):
    """The reference's scalar min/max normalisation (_loaders.py:34-38), float32.  (A mesh whose first material never got
    a face has no rows: nothing to do; the reference's min() would raise on it.)"""
    if rows.shape[0] == 0:
        return
    v_min = rows[:, 0:3].min()
    v_max = rows[:, 0:3].max()
    v_size = v_max - v_min
    max_dim = v_size.max()
    rows[:, 0:3] = (rows[:, 0:3] - v_min) / max_dim - v_size * 0.5 / max_dim


def load_obj(path):
    """Returns [(Mesh, None), ...] like the reference: soup MeshVertex buffer + an (unfilled, zero) index
    buffer of len(faces)*3 `int` entries (_loaders.py:17 never writes it; tutorials pass index_buffer=None)."""
    import ctypes
    from .. import _native
    L = _native.lib()
    handle = ctypes.c_uint64(0)
    if L.rt_obj_load(os.fsencode(path), ctypes.byref(handle)) != 0:
        raise Exception(f"load_obj({path!r}): {L.rt_last_error().decode(errors='replace')}")
    objs = []
    try:
        for i in range(L.rt_obj_mesh_count(handle)):
            nv, nf, fmt = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int(0)
            _native.call("rt_obj_mesh_info", handle, i, ctypes.byref(nv), ctypes.byref(nf), ctypes.byref(fmt))
            mesh_vertices = create_buffer(nv.value, MeshVertex)
            mesh_indices = create_buffer(nf.value * 3, int)
            with mapped(mesh_vertices) as map:
                rows = map.view(np.float32).reshape(nv.value, MeshVertex.itemsize // 4)
                if L.rt_obj_mesh_rows(handle, i, rows.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), rows.shape[1]) != 0:
                    raise Exception(f"load_obj({path!r}): {L.rt_last_error().decode(errors='replace')}")
                _normalise_positions(rows)
            objs.append((Mesh(mesh_vertices, mesh_indices), None))  # mesh + material
    finally:
        L.rt_obj_free(handle)
    return objs


def _load_obj_python(path):
    """load_obj on the Python parser: the definition the native path is tested against."""
    pos, nrm, tex, meshes = _parse_obj(path)
    objs = []
    for m in meshes:
        if not m.materials:
            continue
        mat = m.materials[0]
        c = np.asarray(mat.corners, dtype=np.int64).reshape(-1, 3)
        vertex_count = c.shape[0]
        mesh_vertices = create_buffer(vertex_count, MeshVertex)
        mesh_indices = create_buffer(m.n_faces * 3, int)
        with mapped(mesh_vertices) as map:
            rows = map.view(np.float32).reshape(vertex_count, MeshVertex.itemsize // 4)
            for att in (mat.vertex_format or 'V3F').split('_'):     # a material without faces has no format yet
                if att == 'N3F':
                    rows[:, 4:7] = nrm[c[:, 2]]
                elif att == 'V3F':
                    rows[:, 0:3] = pos[c[:, 0]]
                elif att == 'T2F':
                    rows[:, 8:10] = tex[c[:, 1]]
                else:
                    raise Exception(f'Vertex format in obj {mat.vertex_format} is not supported')
            _normalise_positions(rows)
        objs.append((Mesh(mesh_vertices, mesh_indices), None))  # mesh + material
    return objs
