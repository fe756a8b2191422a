"""ctypes front end for the two CPU checkers (oracle_api.h).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  `Oracle(kind="ref")` drives the UNMODIFIED reference compiled as
oracle/_ref/libfembrain_ref.so; `Oracle(kind="port")` drives the plain-C restatement
oracle/libfembrain_port.so.  Both expose the same methods.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {"ref": os.path.join(_HERE, "_ref", "libfembrain_ref.so"), "port": os.path.join(_HERE, "libfembrain_port.so"),
         "dropin": os.path.join(_HERE, "_ref", "libfembrain_dropin.so")}
_PREFIX = {"ref": "fbref_", "port": "fbport_"}
_loaded: dict = {}


def build(kind: str, reference_root: str = "/root/reference") -> bool:
    """Build one checker with oracle/Makefile.  'ref' needs the reference tree (this container only)."""
    if kind in ("ref", "dropin") and not os.path.isdir(reference_root):
        return os.path.exists(_LIBS[kind])
    subprocess.run(["make", "-C", _HERE, kind, f"REF={reference_root}", "-j8"], check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(_LIBS[kind])


def available(kind: str) -> bool:
    return os.path.exists(_LIBS[kind])


def _lib(kind: str):
    if kind in _loaded:
        return _loaded[kind]
    if not os.path.exists(_LIBS[kind]):
        raise FileNotFoundError(f"{_LIBS[kind]} missing: run `make -C oracle {kind}`")
    lib = C.CDLL(_LIBS[kind])
    p = _PREFIX[kind]
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    sig = {
        "create": (vp, [ci, vp, ci, vp, cd, cd, cd, ci, vp, cd, cd, cd]),
        "destroy": (None, [vp]),
        "r": (ci, [vp]), "nnz_K": (ci, [vp]), "nnz_M": (ci, [vp]), "rows_sys": (ci, [vp]), "nnz_sys": (ci, [vp]),
        "K_csr": (None, [vp, vp, vp, vp]), "M_csr": (None, [vp, vp, vp, vp]), "sys_csr": (None, [vp, vp, vp, vp]),
        "element_maps": (None, [vp, vp, vp]), "element_data": (None, [vp, vp, vp]),
        "super_maps": (None, [vp, vp, vp]), "submatrix_map": (None, [vp, vp]),
        "force_and_matrix": (None, [vp, vp, vp, vp]),
        "set_state": (None, [vp, vp, vp]), "get_state": (None, [vp, vp, vp, vp]),
        "set_external_forces": (None, [vp, vp]), "do_timestep": (ci, [vp]),
        "K_values": (None, [vp, vp]), "rhs": (None, [vp, vp]), "internal_forces": (None, [vp, vp]), "qdelta": (None, [vp, vp]),
        "solve": (ci, [vp, vp, vp, cd, ci]), "solve_iters": (ci, [vp, vp, vp, ci]), "sys_spmv": (None, [vp, vp, vp]), "assign_system": (None, [vp]),
        "assembly_time": (cd, [vp]), "solve_time": (cd, [vp]),
        "polar": (cd, [vp, vp, vp, cd]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, p + name)
        fn.restype, fn.argtypes = res, args
    if kind == "port":
        lib.fbport_create_materials.restype = vp
        lib.fbport_create_materials.argtypes = [ci, vp, ci, vp, vp, vp, vp, ci, vp, cd, cd, cd]
        lib.fbport_deformable_timestep.restype = ci
        lib.fbport_deformable_timestep.argtypes = [vp, ci, ci, vp, vp, ci, ci, ci, vp, ci, ci, cd, C.POINTER(ci)]
        lib.fbport_get_external_forces.restype = None
        lib.fbport_get_external_forces.argtypes = [vp, vp]
    if kind == "ref":
        lib.fbref_save_veg.restype = ci
        lib.fbref_save_veg.argtypes = [vp, C.c_char_p]
        lib.fbref_create_from_veg.restype = vp
        lib.fbref_create_from_veg.argtypes = [C.c_char_p, ci, vp, cd, cd, cd]
        lib.fbref_num_vertices.restype = ci
        lib.fbref_num_vertices.argtypes = [vp]
        lib.fbref_num_tets.restype = ci
        lib.fbref_num_tets.argtypes = [vp]
        lib.fbref_mesh.restype = None
        lib.fbref_mesh.argtypes = [vp, vp, vp, vp, vp, vp]
        if hasattr(lib, "fbref_mt_assembly_seconds"):
            lib.fbref_mt_assembly_seconds.restype = cd
            lib.fbref_mt_assembly_seconds.argtypes = [vp, vp, ci, ci, vp]
        if hasattr(lib, "fbref_force_and_matrix_warp"):   # (a library built before the warp entry was added lacks it)
            lib.fbref_force_and_matrix_warp.restype = None
            lib.fbref_force_and_matrix_warp.argtypes = [vp, vp, ci, vp, vp]
    _loaded[kind] = lib
    return lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def polar(F, tol=1e-6, kind="ref"):
    lib = _lib(kind)
    F = _f64(F).reshape(9)
    R, S = np.zeros(9), np.zeros(9)
    det = getattr(lib, _PREFIX[kind] + "polar")(F.ctypes.data, R.ctypes.data, S.ctypes.data, tol)
    return R.reshape(3, 3), S.reshape(3, 3), det


class Oracle:
    """One deformable model on the CPU checker; mirrors the methods of fembrain_b200.Simulation."""

    def __init__(self, verts=None, tets=None, fixed_verts=(), E=1e7, nu=0.46, rho=1000.0, h=0.0333, damp_mass=0.0,
                 damp_stiffness=0.01, kind="ref", materials=None, veg_path=None):
        """materials = (E[nT], nu[nT], rho[nT]) arrays (port only); veg_path = let the reference's own .veg loader
        build the mesh and its materials (ref only)."""
        self.kind = kind
        self._lib = _lib(kind)
        self._p = _PREFIX[kind]
        fx = _i32(fixed_verts)
        if veg_path is not None:
            assert kind == "ref", "only the compiled reference has the .veg loader"
            self._h = self._lib.fbref_create_from_veg(str(veg_path).encode(), len(fx), fx.ctypes.data if len(fx) else None, h,
                                                      damp_mass, damp_stiffness)
            if not self._h:
                raise RuntimeError("reference could not load " + str(veg_path))
            self.nV, self.nT = self._lib.fbref_num_vertices(self._h), self._lib.fbref_num_tets(self._h)
        else:
            v, t = _f64(verts), _i32(tets)
            self.nV, self.nT = len(v), len(t)
            if materials is not None:
                assert kind == "port", "per-element arrays: port only (the reference reads them from a .veg)"
                me, mn, mr = (_f64(m) for m in materials)
                self._h = self._lib.fbport_create_materials(self.nV, v.ctypes.data, self.nT, t.ctypes.data, me.ctypes.data, mn.ctypes.data,
                                                            mr.ctypes.data, len(fx), fx.ctypes.data if len(fx) else None, h, damp_mass,
                                                            damp_stiffness)
            else:
                self._h = self._fn("create")(self.nV, v.ctypes.data, self.nT, t.ctypes.data, E, nu, rho, len(fx),
                                             fx.ctypes.data if len(fx) else None, h, damp_mass, damp_stiffness)
        if not self._h:
            raise RuntimeError("oracle create failed")
        self.r = self._fn("r")(self._h)
        self.nnz_K = self._fn("nnz_K")(self._h)
        self.nnz_M = self._fn("nnz_M")(self._h)
        self.rows_sys = self._fn("rows_sys")(self._h)
        self.nnz_sys = self._fn("nnz_sys")(self._h)

    def _fn(self, name):
        return getattr(self._lib, self._p + name)

    def close(self):
        if getattr(self, "_h", None):
            self._fn("destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def save_veg(self, path):
        """The reference's own writer (TetMesh::save) on the mesh it loaded (ref only)."""
        assert self.kind == "ref"
        return self._lib.fbref_save_veg(self._h, str(path).encode())

    def mesh(self):
        """(verts, tets, E, nu, rho) as the reference holds them after loading a .veg (ref only)."""
        assert self.kind == "ref"
        v, t = np.zeros((self.nV, 3)), np.zeros((self.nT, 4), np.int32)
        E, nu, rho = np.zeros(self.nT), np.zeros(self.nT), np.zeros(self.nT)
        self._lib.fbref_mesh(self._h, v.ctypes.data, t.ctypes.data, E.ctypes.data, nu.ctypes.data, rho.ctypes.data)
        return v, t, E, nu, rho

    def _csr(self, name, n, nnz, values=True):
        ia, ja = np.zeros(n + 1, np.int32), np.zeros(max(nnz, 1), np.int32)[:nnz]
        a = np.zeros(max(nnz, 1))[:nnz] if values else None
        self._fn(name)(self._h, ia.ctypes.data, ja.ctypes.data, a.ctypes.data if values else None)
        return ia, ja, a

    def K_csr(self, values=True):
        return self._csr("K_csr", self.r, self.nnz_K, values)

    def M_csr(self):
        return self._csr("M_csr", self.r, self.nnz_M)

    def sys_csr(self, values=True):
        return self._csr("sys_csr", self.rows_sys, self.nnz_sys, values)

    def element_maps(self):
        row, col = np.zeros(4 * self.nT, np.int32), np.zeros(16 * self.nT, np.int32)
        self._fn("element_maps")(self._h, row.ctypes.data, col.ctypes.data)
        return row.reshape(-1, 4), col.reshape(-1, 16)

    def element_data(self):
        mi, k0 = np.zeros(16 * self.nT), np.zeros(144 * self.nT)
        self._fn("element_data")(self._h, mi.ctypes.data, k0.ctypes.data)
        return mi.reshape(-1, 16), k0.reshape(-1, 144)

    def super_maps(self):
        sr, si = np.zeros(max(self.rows_sys, 1), np.int32)[: self.rows_sys], np.zeros(max(self.nnz_sys, 1), np.int32)[: self.nnz_sys]
        self._fn("super_maps")(self._h, sr.ctypes.data, si.ctypes.data)
        return sr, si

    def submatrix_map(self):
        idx = np.zeros(max(self.nnz_M, 1), np.int32)[: self.nnz_M]
        self._fn("submatrix_map")(self._h, idx.ctypes.data)
        return idx

    def force_and_matrix(self, u):
        u = _f64(u).reshape(-1)
        f, a = np.zeros(self.r), np.zeros(self.nnz_K)
        self._fn("force_and_matrix")(self._h, u.ctypes.data, f.ctypes.data, a.ctypes.data)
        return f, a

    def force_and_matrix_warp(self, u, warp):
        """ComputeForceAndStiffnessMatrix(u, f, K, warp) of the compiled reference (kind "ref" only: the port restates warp = 1)."""
        if self.kind != "ref":
            raise NotImplementedError("warp modes are checked against the compiled reference")
        u = _f64(u).reshape(-1)
        f, a = np.zeros(self.r), np.zeros(self.nnz_K)
        fn = self._fn("force_and_matrix_warp")
        fn(self._h, u.ctypes.data, int(warp), f.ctypes.data, a.ctypes.data)
        return f, a

    def mt_assembly_seconds(self, u, threads, reps=2):
        """CorotationalLinearFEMMT (the reference's pthread assembly) on this mesh: (best wall-clock seconds per
        ComputeForceAndStiffnessMatrix, max |f_mt - f_single|).  kind "ref" only."""
        u = _f64(u).reshape(-1)
        diff = C.c_double(0)
        sec = self._lib.fbref_mt_assembly_seconds(self._h, u.ctypes.data, int(threads), int(reps), C.addressof(diff))
        return float(sec), float(diff.value)

    def set_state(self, q, qvel=None):
        q = _f64(q).reshape(-1)
        qv = _f64(qvel).reshape(-1) if qvel is not None else None
        self._fn("set_state")(self._h, q.ctypes.data, qv.ctypes.data if qv is not None else None)

    def get_state(self):
        q, qv, qa = np.zeros(self.r), np.zeros(self.r), np.zeros(self.r)
        self._fn("get_state")(self._h, q.ctypes.data, qv.ctypes.data, qa.ctypes.data)
        return q, qv, qa

    def set_external_forces(self, f):
        f = _f64(f).reshape(-1)
        assert f.size == self.r
        self._fn("set_external_forces")(self._h, f.ctypes.data)

    def do_timestep(self):
        return self._fn("do_timestep")(self._h)

    def _vec(self, name, n):
        out = np.zeros(max(n, 1))[:n]
        self._fn(name)(self._h, out.ctypes.data)
        return out

    def K_values(self):
        return self._vec("K_values", self.nnz_K)

    def rhs(self):
        return self._vec("rhs", self.rows_sys)

    def internal_forces(self):
        return self._vec("internal_forces", self.r)

    def qdelta(self):
        return self._vec("qdelta", self.r)

    def solve(self, b=None, eps=1e-6, max_iter=10000):
        x = np.zeros(self.rows_sys)
        bb = _f64(b) if b is not None else None
        it = self._fn("solve")(self._h, bb.ctypes.data if bb is not None else None, x.ctypes.data, eps, max_iter)
        return x, it

    def solve_iters(self, iters, b=None):
        x = np.zeros(self.rows_sys)
        bb = _f64(b) if b is not None else None
        it = self._fn("solve_iters")(self._h, bb.ctypes.data if bb is not None else None, x.ctypes.data, iters)
        return x, it

    def load_system_from_K(self):
        self._fn("assign_system")(self._h)

    def sys_spmv(self, x):
        x = _f64(x)
        y = np.zeros(self.rows_sys)
        self._fn("sys_spmv")(self._h, x.ctypes.data, y.ctypes.data)
        return y

    # -- Deformable::timestep restatement (port only; see oracle/deformable_port.inc) ---------------------------
    def deformable_timestep(self, gravity=False, haptic_idx=(), haptic_forces=(), in_progress=True, rings=5, edges=None,
                            quirk=False, floor=False, floor_y=0.0, contacts=0):
        assert self.kind == "port", "Deformable.cpp cannot be compiled here; only the restatement has this call"
        idx, f = _i32(haptic_idx), _f64(haptic_forces).reshape(-1)
        e = _i32(edges).reshape(-1) if edges is not None else None
        ct = C.c_int(contacts)
        rc = self._lib.fbport_deformable_timestep(self._h, int(gravity), len(idx), idx.ctypes.data if len(idx) else None,
                                                  f.ctypes.data if f.size else None, int(in_progress), rings,
                                                  (e.size // 2) if e is not None else 0, e.ctypes.data if e is not None else None,
                                                  int(quirk), int(floor), floor_y, C.byref(ct))
        return rc, ct.value

    def get_external_forces(self):
        f = np.zeros(self.r)
        self._lib.fbport_get_external_forces(self._h, f.ctypes.data)
        return f

    def assembly_time(self):
        return self._fn("assembly_time")(self._h)

    def solve_time(self):
        return self._fn("solve_time")(self._h)


# -- the reference's own `class Deformable`, compiled (oracle/deformable_harness.cpp) ------------------------------------------
def _def_lib(which="ref"):
    """which = "ref": the compiled reference (fbdef_*); "dropin": the reference's Deformable.cpp compiled with the INTEGRATION.md
    substitutions on top of libfembrain_b200.so (fbdrop_*, oracle/_ref/libfembrain_dropin.so; needs a GPU to create anything)."""
    if which == "dropin":
        if "dropin" not in _loaded:
            if not os.path.exists(_LIBS["dropin"]):
                raise FileNotFoundError(f"{_LIBS['dropin']} missing: run `make -C oracle dropin`")
            _loaded["dropin"] = C.CDLL(_LIBS["dropin"])
        lib = _loaded["dropin"]
    else:
        lib = _lib("ref")
    if getattr(lib, "_fbdef_bound", False):
        return lib
    vp, ci, cd, cf = C.c_void_p, C.c_int, C.c_double, C.c_float
    pre = "fbdrop_" if which == "dropin" else "fbdef_"
    sig = {
        "fbdef_create": (vp, [ci, vp, ci, vp, ci, vp]), "fbdef_destroy": (None, [vp]),
        "fbdef_num_vertices": (ci, [vp]), "fbdef_num_cells": (ci, [vp]), "fbdef_num_edges": (ci, [vp]),
        "fbdef_mesh": (None, [vp, vp, vp, vp]), "fbdef_node_neighbors": (ci, [vp, ci, ci, vp]),
        "fbdef_set_gravity": (None, [vp, ci]), "fbdef_set_haptic_radius": (None, [vp, ci]), "fbdef_set_floor": (None, [vp, cf]),
        "fbdef_floor_y": (cf, [vp]), "fbdef_set_haptic": (None, [vp, ci, vp, vp, ci]), "fbdef_timestep": (None, [vp]),
        "fbdef_get_state": (None, [vp, vp, vp, vp]), "fbdef_set_state": (None, [vp, vp, vp]),
        "fbdef_get_external_forces": (None, [vp, vp]), "fbdef_contacts": (ci, [vp]), "fbdef_set_contacts": (None, [vp, ci]),
        "fbdef_positions": (None, [vp, vp]), "fbdef_aabb": (None, [vp, vp, vp]),
        "fbdef_pick_vertices": (ci, [vp, vp, vp, ci, vp, vp]), "fbdef_pick_vertex": (ci, [vp, vp, vp, vp]),
        "fbdef_sample_mesh": (ci, [ci, ci, ci, ci, cd, cd, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, pre + name[len("fbdef_"):])
        fn.restype, fn.argtypes = res, args
        setattr(lib, name, fn)   # both libraries are driven through the fbdef_* names below
    lib._fbdef_bound = True
    return lib


def dropin_available() -> bool:
    return os.path.exists(_LIBS["dropin"])


def _libc_fflush():
    try:
        C.CDLL(None).fflush(None)
    except Exception:
        pass


def apply_fem_displacements(rest_xyzw, displacements):
    """GPUPoly::applyFemDisplacements (src/implicit/OclPolygonizer.cpp:1543-1584) with the reference's own OpenCL kernel
    ApplyVertexDeformations (data/opencl/Polygonizer.cl:1417-1427) compiled from its text for the CPU
    (oracle/cl_kernel_harness.cpp): float4 out = rest + float4(displacement, 0)."""
    lib = _lib("ref")
    fn = lib.fbcl_apply_fem_displacements
    fn.restype, fn.argtypes = None, [C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
    rest = np.ascontiguousarray(rest_xyzw, dtype=np.float32).reshape(-1, 4)
    d = _f64(displacements).reshape(-1)
    n = len(rest)
    assert d.size >= 3 * n
    out = np.full((n + 64, 4), np.float32(-7.0))   # work items past ctVertices must not write
    fn(n, rest.ctypes.data, d.ctypes.data, out.ctypes.data)
    assert np.all(out[n:] == np.float32(-7.0))
    return out[:n].copy()


def cl_kernel_available() -> bool:
    return available("ref") and hasattr(_lib("ref"), "fbcl_apply_fem_displacements")


def deformable_available() -> bool:
    """True when oracle/_ref holds the compiled Deformable (built from /root/reference by oracle/Makefile)."""
    if not available("ref"):
        return False
    try:
        _def_lib()
        return True
    except AttributeError:
        return False


def sample_mesh(which, a=0, b=0, c=0, x=0.0, y=0.0):
    """The reference's own VolMeshSamples generators (src/deformable/VolMeshSamples.cpp:15-253), compiled:
    'one_tetra', 'two_tetra', 'truth_cube' (a, b, c = nx, ny, nz; x = cellsize), 'egg_shell' (a, b = hseg, vseg; x radius, y thickness)."""
    lib = _def_lib()
    code = {"one_tetra": 0, "two_tetra": 1, "truth_cube": 2, "egg_shell": 3}[which]
    nV, nT = C.c_int(0), C.c_int(0)
    if lib.fbdef_sample_mesh(code, a, b, c, x, y, C.byref(nV), C.byref(nT), None, None) != 0:
        raise RuntimeError("sample mesh failed")
    v, t = np.zeros((nV.value, 3)), np.zeros((nT.value, 4), np.int32)
    lib.fbdef_sample_mesh(code, a, b, c, x, y, C.byref(nV), C.byref(nT), v.ctypes.data, t.ctypes.data)
    return v, t


class RefDeformable:
    """The UNMODIFIED reference `Deformable` (Deformable.cpp, CuttableMesh / VolMesh under it) on a tet mesh: the compiled
    checker for Deformable::timestep, applyHapticForces, the floor post-step, pickVertices / pickVertex and the
    get_node_neighbors quirk.  Material, time step, damping and CG settings are the ones Deformable hard-codes."""

    def __init__(self, verts, tets, fixed_verts=(), lib="ref"):
        """lib="dropin": the same reference class compiled with its integrator and force model replaced by the CUDA-backed
        subclasses of include/fembrain_b200_vega_classes.hpp (oracle/stubs/prelude_dropin.h)."""
        self._lib = _def_lib(lib)
        v, t, fx = _f64(verts).reshape(-1, 3), _i32(tets).reshape(-1, 4), _i32(fixed_verts)
        # CuttableMesh's setup runs the reference's mesh self-tests, which print to stdout (test_VolMesh.cpp): keep them off it
        sys.stdout.flush()
        saved, null = os.dup(1), os.open(os.devnull, os.O_WRONLY)
        os.dup2(null, 1)
        try:
            self._h = self._lib.fbdef_create(len(v), v.ctypes.data, len(t), t.ctypes.data, len(fx), fx.ctypes.data if len(fx) else None)
            _libc_fflush()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(null)
        if not self._h:
            raise RuntimeError("reference VolMesh::setup failed" if lib == "ref" else "drop-in Deformable could not be created (no usable GPU?)")
        self.nV, self.nT = self._lib.fbdef_num_vertices(self._h), self._lib.fbdef_num_cells(self._h)
        self.r = 3 * self.nV

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fbdef_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def mesh(self):
        """(rest positions, cells, edges [nE,2] in VolMesh::m_vEdges order) of the CuttableMesh Deformable holds."""
        nE = self._lib.fbdef_num_edges(self._h)
        x, c, e = np.zeros((self.nV, 3)), np.zeros((self.nT, 4), np.int32), np.zeros((max(nE, 1), 2), np.int32)
        self._lib.fbdef_mesh(self._h, x.ctypes.data, c.ctypes.data, e.ctypes.data)
        return x, c, e[:nE]

    def node_neighbors(self, v):
        out = np.zeros(4096, np.int32)
        n = self._lib.fbdef_node_neighbors(self._h, int(v), len(out), out.ctypes.data)
        return out[:n].copy()

    def set_gravity(self, on):
        self._lib.fbdef_set_gravity(self._h, int(on))

    def set_haptic_radius(self, rings):
        self._lib.fbdef_set_haptic_radius(self._h, int(rings))

    def set_floor(self, y):
        self._lib.fbdef_set_floor(self._h, float(y))
        return float(self._lib.fbdef_floor_y(self._h))

    def set_haptic(self, indices, forces, in_progress=True):
        idx, f = _i32(indices), _f64(forces).reshape(-1)
        self._lib.fbdef_set_haptic(self._h, len(idx), idx.ctypes.data if len(idx) else None, f.ctypes.data if f.size else None, int(in_progress))

    def timestep(self):
        self._lib.fbdef_timestep(self._h)

    def get_state(self):
        q, qv, qa = np.zeros(self.r), np.zeros(self.r), np.zeros(self.r)
        self._lib.fbdef_get_state(self._h, q.ctypes.data, qv.ctypes.data, qa.ctypes.data)
        return q, qv, qa

    def set_state(self, q, qvel):
        q, qv = _f64(q).reshape(-1), _f64(qvel).reshape(-1)
        self._lib.fbdef_set_state(self._h, q.ctypes.data, qv.ctypes.data)

    def external_forces(self):
        f = np.zeros(self.r)
        self._lib.fbdef_get_external_forces(self._h, f.ctypes.data)
        return f

    @property
    def contacts(self):
        return self._lib.fbdef_contacts(self._h)

    def set_contacts(self, n):
        self._lib.fbdef_set_contacts(self._h, int(n))

    def positions(self):
        x = np.zeros((self.nV, 3))
        self._lib.fbdef_positions(self._h, x.ctypes.data)
        return x

    def aabb(self):
        lo, hi = np.zeros(3, np.float32), np.zeros(3, np.float32)
        self._lib.fbdef_aabb(self._h, lo.ctypes.data, hi.ctypes.data)
        return lo, hi

    def pick_vertices(self, lo, hi):
        lo, hi = _f64(lo), _f64(hi)
        idx, co = np.zeros(max(self.nV, 1), np.int32), np.zeros(3 * max(self.nV, 1))
        n = self._lib.fbdef_pick_vertices(self._h, lo.ctypes.data, hi.ctypes.data, self.nV, idx.ctypes.data, co.ctypes.data)
        return idx[:n].copy(), co[:3 * n].reshape(-1, 3).copy()

    def pick_vertex(self, world_pos):
        w, d, v = _f64(world_pos), C.c_double(0), np.zeros(3)
        i = self._lib.fbdef_pick_vertex(self._h, w.ctypes.data, C.byref(d), v.ctypes.data)
        return i, d.value, v


# -- Deformable's vertex queries (force producers), restated with numpy.  PINNED against the compiled
#    Deformable::pickVertices / CuttableMesh::findClosestVertex (RefDeformable above; tests/test_oracle_deformable.py). ------
def pick_vertices(rest_pos, u, box_lo, box_hi):
    """Deformable::pickVertices (src/deformable/Deformable.cpp:430-448) with Contains<double> (src/graphics/AABB.h:84-92, closed
    box) on pos = restpos + u (VolMesh::displace, src/deformable/VolMesh.cpp:1370-1385): ascending indices and coordinates."""
    pos = _f64(rest_pos).reshape(-1, 3) + _f64(u).reshape(-1, 3)
    lo, hi = _f64(box_lo), _f64(box_hi)
    inside = np.all((pos >= lo) & (pos <= hi), axis=1)
    idx = np.nonzero(inside)[0].astype(np.int32)
    return idx, pos[idx]


def pick_vertex(rest_pos, u, world_pos):
    """Deformable::pickVertex -> CuttableMesh::findClosestVertex (src/deformable/Deformable.cpp:422-428,
    src/deformable/CuttableMesh.cpp:511-526): strict '<' on dx*dx + dy*dy + dz*dz in index order = first minimum."""
    pos = _f64(rest_pos).reshape(-1, 3) + _f64(u).reshape(-1, 3)
    if len(pos) == 0:
        return -1, float(np.sqrt(np.finfo(np.float64).max)), None
    d = _f64(world_pos) - pos
    d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
    i = int(np.argmin(d2))  # first occurrence of the minimum
    return i, float(np.sqrt(d2[i])), pos[i]
