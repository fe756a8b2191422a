"""ctypes binding of include/fembrain_b200.h — thin plumbing over the C ABI, no compute here.

`Simulation` mirrors the reference's integrator surface (SetExternalForces / DoTimestep / GetqState,
vegafem/integrator/integratorBase.h:107-205) plus the parity-inspection hooks.  There is no CPU path:
importing works without a GPU (so symbols can be checked), creating a Simulation without one raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FEMBRAIN_B200_LIB") or os.path.join(_HERE, "libfembrain_b200.so")   # (override: A/B timing of two builds)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fembrain_b200.h")

FB_OK = 0
FB_ERR_INVALID_ARGUMENT = 1
FB_ERR_NO_DEVICE = 2
FB_ERR_CUDA = 3
FB_ERR_OUT_OF_MEMORY = 4
FB_ERR_SOLVER_NOT_CONVERGED = 5
FB_ERR_BAD_MESH = 6
FB_ERR_COMM = 7
FB_ERR_NOT_SUPPORTED = 8


class FbParams(C.Structure):
    _fields_ = [
        ("youngs_modulus", C.c_double),
        ("poisson_ratio", C.c_double),
        ("density", C.c_double),
        ("timestep", C.c_double),
        ("damping_mass", C.c_double),
        ("damping_stiffness", C.c_double),
        ("cg_epsilon", C.c_double),
        ("cg_max_iterations", C.c_int),
        ("polar_tolerance", C.c_double),
        ("internal_force_scaling", C.c_double),
        ("device", C.c_int),
        ("keep_raw_stiffness", C.c_int),
        ("solver_variant", C.c_int),
        ("warm_start", C.c_int),
        ("reserved", C.c_int * 4),
    ]


class FemBrainError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str):
        super().__init__(f"{where}: status {status} ({detail})")
        self.status = status


_lib = None


def step_many(sims, host_threads=8):
    """fb_step_many: one DoTimestep of every Simulation in `sims` (independent contexts), issued from `host_threads` native threads.
    Returns the list of per-context status codes; raises FemBrainError for the first failure."""
    lib = load_library()
    n = len(sims)
    handles = (C.c_void_p * n)(*[s._h for s in sims])
    status = (C.c_int * n)()
    st = lib.fb_step_many(handles, n, int(host_threads), status)
    if st != 0:
        raise FemBrainError(st, "fb_step_many", lib.fb_last_error_string().decode(errors="replace"))
    return list(status)


def _preload_bundled_nccl():
    """The library's NEEDED libnccl.so.2 resolves to whichever copy the process maps first.  PyTorch ships its own (newer) NCCL
    and fails to import (`undefined symbol: ncclDevCommCreate`) when an older system copy is already mapped — which is what
    happens if this module is imported BEFORE torch.  Mapping the wheel's copy first makes the order irrelevant; without the
    wheel the system library is used as linked."""
    import importlib.util

    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    if spec is None or not spec.submodule_search_locations:
        return
    for root in spec.submodule_search_locations:
        cand = os.path.join(root, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
            return


def load_library():
    """Load libfembrain_b200.so; raise loudly if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m fembrain_b200.build` (nvcc, sm_100a). "
            "fembrain_b200 has no CPU or PyTorch fallback."
        )
    _preload_bundled_nccl()
    lib = C.CDLL(LIB_PATH)
    vp, ci, cd, ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
    pp = C.POINTER(vp)
    prm = C.POINTER(FbParams)
    sig = {
        "fb_default_params": (None, [prm]),
        "fb_abi_version": (ci, []),
        "fb_status_string": (C.c_char_p, [ci]),
        "fb_last_error_string": (C.c_char_p, []),
        "fb_create": (ci, [pp, ci, vp, ci, vp, ci, vp, prm]),
        "fb_create_with_materials": (ci, [pp, ci, vp, ci, vp, ci, vp, vp, vp, vp, prm]),
        "fb_create_with_constrained_dofs": (ci, [pp, ci, vp, ci, vp, ci, vp, prm]),
        "fb_destroy": (None, [vp]),
        "fb_sync_force_model": (ci, [vp, ci, vp, ci, vp, ci, vp]),
        "fb_set_fixed_vertices": (ci, [vp, ci, vp]),
        "fb_num_vertices": (ci, [vp]), "fb_num_tets": (ci, [vp]), "fb_num_dofs": (ci, [vp]),
        "fb_num_constrained_dofs": (ci, [vp]), "fb_num_local_dofs": (ci, [vp]),
        "fb_nnz_stiffness": (ll, [vp]), "fb_nnz_mass": (ll, [vp]), "fb_nnz_system": (ll, [vp]),
        "fb_set_external_forces": (ci, [vp, vp]), "fb_add_external_forces": (ci, [vp, vp]),
        "fb_set_external_forces_to_zero": (ci, [vp]), "fb_get_external_forces": (ci, [vp, vp]),
        "fb_set_state": (ci, [vp, vp, vp, vp]), "fb_get_state": (ci, [vp, vp, vp, vp]),
        "fb_reset_to_rest": (ci, [vp]),
        "fb_set_external_forces_dev": (ci, [vp, vp]), "fb_get_state_dev": (ci, [vp, vp, vp, vp]),
        "fb_displacements_dev": (vp, [vp]),
        "fb_set_warp": (ci, [vp, ci]), "fb_get_warp": (ci, [vp]),
        "fb_step_many": (ci, [vp, ci, ci, vp]),
        "fb_set_timestep": (ci, [vp, cd]), "fb_set_damping": (ci, [vp, cd, cd]),
        "fb_set_internal_force_scaling": (ci, [vp, cd]), "fb_set_cg": (ci, [vp, cd, ci]),
        "fb_set_grid": (ci, [vp, ci, ci, ci]), "fb_set_solver": (ci, [vp, ci, ci]),
        "fb_get_solver": (ci, [vp, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]), "fb_solver_name": (C.c_char_p, [ci]),
        "fb_get_solver_levels": (ci, [vp, ci, vp, vp]),
        "fb_get_solver_smoother": (ci, [vp, vp, vp, vp, vp]),
        "fb_step": (ci, [vp]),
        "fb_deformable_timestep": (ci, [vp]), "fb_deformable_set_gravity": (ci, [vp, ci]),
        "fb_deformable_set_floor": (ci, [vp, ci, cd]),
        "fb_deformable_set_haptic_forces": (ci, [vp, ci, vp, vp, ci]),
        "fb_deformable_set_haptic_neighborhood": (ci, [vp, ci]),
        "fb_deformable_contact_count": (ci, [vp]),
        "fb_deformable_set_edge_list": (ci, [vp, ci, vp, ci]),
        "fb_deformable_pick_vertices": (ci, [vp, vp, vp, ci, vp, vp, vp]), "fb_deformable_pick_vertex": (ci, [vp, vp, vp, vp, vp]),
        "fb_force_assembly_seconds": (cd, [vp]), "fb_system_solve_seconds": (cd, [vp]), "fb_step_seconds": (cd, [vp]),
        "fb_last_cg_iterations": (ci, [vp]), "fb_last_cg_residual_ratio": (cd, [vp]),
        "fb_veg_save": (ci, [C.c_char_p, ci, ci, vp, ci, vp, vp, vp, vp]),
        "fb_create_batch": (ci, [pp, ci, vp, vp, vp, vp, vp, vp, prm]), "fb_batch_count": (ci, [vp]),
        "fb_batch_offsets": (ci, [vp, vp, vp]), "fb_batch_last_cg_iterations": (ci, [vp, vp, vp]),
        "fb_partition_ordering": (ci, [ci, ci, vp, ci, vp, C.POINTER(ci)]), "fb_partition_reordered": (ci, [vp]),
        "fb_kernel_launches": (ll, [vp]), "fb_device_bytes": (C.c_size_t, [vp]), "fb_trim_memory": (ci, []), "fb_experiments_built": (ci, []),
        "fb_check_guards": (ci, [C.POINTER(ll), C.POINTER(ll)]),
        "fb_get_stiffness_csr": (ci, [vp, vp, vp]), "fb_get_mass_csr": (ci, [vp, vp, vp, vp]),
        "fb_get_system_csr": (ci, [vp, vp, vp, vp]), "fb_get_element_maps": (ci, [vp, vp, vp]),
        "fb_get_element_data": (ci, [vp, vp, vp]), "fb_get_super_maps": (ci, [vp, vp, vp]),
        "fb_get_submatrix_map": (ci, [vp, vp]), "fb_get_constrained_dofs": (ci, [vp, vp]),
        "fb_compute_force_and_matrix": (ci, [vp, vp, vp, vp]),
        "fb_get_effective_stiffness_values": (ci, [vp, vp]), "fb_get_rhs": (ci, [vp, vp]),
        "fb_get_internal_forces": (ci, [vp, vp]), "fb_get_qdelta": (ci, [vp, vp]),
        "fb_solve": (ci, [vp, vp, vp, cd, ci, C.POINTER(ci)]), "fb_system_multiply": (ci, [vp, vp, vp]),
        "fb_veg_load": (ci, [C.c_char_p, C.POINTER(ci), C.POINTER(ci), pp, pp, pp, pp, pp]),
        "fb_veg_free": (None, [vp]),
        "fb_tetgen_load": (ci, [C.c_char_p, C.POINTER(ci), C.POINTER(ci), pp, pp, pp, pp, pp]),
        "fb_create_from_veg": (ci, [pp, C.c_char_p, ci, vp, prm]),
        "fb_export_positions_float4": (ci, [vp, ci, vp, vp]),
        "fb_export_positions_float4_dev": (ci, [vp, ci, vp, vp]),
        "fb_timer_start": (ci, [vp]), "fb_timer_stop": (ci, [vp, C.POINTER(cd)]),
        "fb_set_profiling": (ci, [vp, ci]),
        "fb_get_spmv_profile": (ci, [vp, C.POINTER(cd), C.POINTER(ci), C.POINTER(cd)]),
        "fb_bench_spmv": (ci, [vp, ci, C.POINTER(cd)]), "fb_bench_assembly": (ci, [vp, ci, C.POINTER(cd)]),
        "fb_bench_cg_iteration": (ci, [vp, ci, C.POINTER(cd)]),
        "fb_comm_unique_id": (ci, [vp]),
        "fb_create_partitioned": (ci, [pp, ci, vp, ci, vp, ci, vp, prm, ci, ci, vp]),
        "fb_partition_range": (ci, [vp, C.POINTER(ci), C.POINTER(ci)]),
        "fb_partition_peer_memory": (ci, [vp]),
        "fb_partition_local_range": (ci, [vp, C.POINTER(ci), C.POINTER(ci)]), "fb_partition_local_to_global": (ci, [vp, vp]),
        "fb_set_external_forces_owned": (ci, [vp, vp]), "fb_get_state_owned": (ci, [vp, vp, vp, vp]),
        "fb_plan_partition": (ci, [ci, ci, vp, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib._fb_signatures = sig
    _lib = lib
    return lib


def default_params(**overrides) -> FbParams:
    lib = load_library()
    p = FbParams()
    lib.fb_default_params(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(f"fb_params has no field {k}")
        setattr(p, k, v)
    return p


def comm_unique_id() -> bytes:
    """ncclUniqueId (128 bytes) for fb_create_partitioned; create on rank 0 and broadcast."""
    lib = load_library()
    buf = (C.c_char * 128)()
    st = lib.fb_comm_unique_id(C.cast(buf, C.c_void_p))
    if st != FB_OK:
        raise FemBrainError(st, "fb_comm_unique_id", lib.fb_last_error_string().decode())
    return bytes(buf)


def plan_partition(num_vertices, tets, world, rank):
    """Host-only partition plan (no GPU): dict with range, local-to-global map, local tets and halo lists."""
    lib = load_library()
    t = np.ascontiguousarray(tets, dtype=np.int32).reshape(-1, 4)
    counts = np.zeros(7, np.int32)
    args = (num_vertices, len(t), t.ctypes.data if len(t) else None, world, rank)
    st = lib.fb_plan_partition(*args, counts.ctypes.data, None, None, None, None, None, None, None)
    if st != FB_OK:
        raise FemBrainError(st, "fb_plan_partition", lib.fb_last_error_string().decode())
    n = [max(int(c), 1) for c in counts]
    l2g, lt = np.zeros(n[2], np.int32), np.zeros(n[3], np.int32)
    nbr, sc, rc = np.zeros(n[4], np.int32), np.zeros(n[4], np.int32), np.zeros(n[4], np.int32)
    sg, rg = np.zeros(n[5], np.int32), np.zeros(n[6], np.int32)
    st = lib.fb_plan_partition(*args, counts.ctypes.data, l2g.ctypes.data, lt.ctypes.data, nbr.ctypes.data, sc.ctypes.data,
                               rc.ctypes.data, sg.ctypes.data, rg.ctypes.data)
    if st != FB_OK:
        raise FemBrainError(st, "fb_plan_partition", lib.fb_last_error_string().decode())
    k = int(counts[4])
    so, ro = np.concatenate([[0], np.cumsum(sc[:k])]), np.concatenate([[0], np.cumsum(rc[:k])])
    return {
        "begin": int(counts[0]), "end": int(counts[1]), "l2g": l2g[: counts[2]], "local_tets": lt[: counts[3]],
        "neighbours": [int(x) for x in nbr[:k]],
        "send": {int(nbr[i]): sg[so[i]:so[i + 1]] for i in range(k)},
        "recv": {int(nbr[i]): rg[ro[i]:ro[i + 1]] for i in range(k)},
    }


def partition_ordering(num_vertices, tets, world):
    """fb_partition_ordering: (order[new] = caller's vertex id, reordered flag) of the numbering the row blocks are cut from."""
    lib = load_library()
    t = _i32(tets).reshape(-1)
    order, flag = np.zeros(max(num_vertices, 1), np.int32), C.c_int(0)
    st = lib.fb_partition_ordering(num_vertices, len(t) // 4, t.ctypes.data if len(t) else None, world, order.ctypes.data, C.byref(flag))
    if st != FB_OK:
        raise FemBrainError(st, "fb_partition_ordering", lib.fb_last_error_string().decode())
    return order[:num_vertices], bool(flag.value)


def experiments_built() -> bool:
    """True when libfembrain_b200.so carries the shelved experiments of csrc/experiments/ (build --experiments)."""
    return bool(load_library().fb_experiments_built())


def check_guards():
    """(allocations checked, allocations with an overwritten guard band) — FEMBRAIN_B200_GUARD=1 processes only."""
    lib = load_library()
    a, b = C.c_longlong(0), C.c_longlong(0)
    st = lib.fb_check_guards(C.byref(a), C.byref(b))
    if st != FB_OK:
        raise FemBrainError(st, "fb_check_guards", lib.fb_last_error_string().decode())
    return a.value, b.value


def trim_memory():
    """Give the device memory that destroyed contexts left in the stream-ordered pool back to the driver."""
    return load_library().fb_trim_memory()


def tetgen_load(basename):
    """fb_tetgen_load: <basename>.node / .ele with the rules of the reference's TetMesh(char*) constructor."""
    return veg_load(basename, _fn="fb_tetgen_load")


def veg_load(path, _fn="fb_veg_load"):
    """fb_veg_load: (verts [nV,3], tets [nT,4] 0-based, E[nT], nu[nT], density[nT]) with the reference loader's rules."""
    lib = load_library()
    nv, nt = C.c_int(0), C.c_int(0)
    ptrs = [C.c_void_p() for _ in range(5)]
    st = getattr(lib, _fn)(str(path).encode(), C.byref(nv), C.byref(nt), *[C.byref(p) for p in ptrs])
    if st != FB_OK:
        raise FemBrainError(st, _fn, lib.fb_last_error_string().decode())
    try:
        def arr(p, ctype, n, dt):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(max(n, 1),))[:n].astype(dt, copy=True)
        v = arr(ptrs[0], C.c_double, 3 * nv.value, np.float64).reshape(-1, 3)
        t = arr(ptrs[1], C.c_int, 4 * nt.value, np.int32).reshape(-1, 4)
        E, nu, rho = (arr(p, C.c_double, nt.value, np.float64) for p in ptrs[2:])
    finally:
        for p in ptrs:
            lib.fb_veg_free(p)
    return v, t, E, nu, rho


VEG_STYLE_FEMBRAIN, VEG_STYLE_VEGA = 0, 1


def veg_save(path, verts, tets, E=None, nu=None, density=None, style=VEG_STYLE_VEGA):
    """fb_veg_save: VolMeshIO::writeVega (style FEMBRAIN) or VolumetricMesh::save (style VEGA) file formats."""
    lib = load_library()
    v, t = _f64(verts).reshape(-1, 3), _i32(tets).reshape(-1, 4)
    mats = [None if a is None else _f64(a).reshape(-1) for a in (E, nu, density)]
    st = lib.fb_veg_save(str(path).encode(), style, len(v), _ptr(v), len(t), _ptr(t), *[None if a is None else _ptr(a) for a in mats])
    if st != FB_OK:
        raise FemBrainError(st, "fb_veg_save", lib.fb_last_error_string().decode())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


class Simulation:
    """One deformable model on one B200: the C-ABI context behind the reference's integrator interface."""

    def __init__(self, verts=None, tets=None, fixed_verts=(), constrained_dofs=None, materials=None, partition=None, veg_path=None,
                 batch=None, **params):
        """batch = [(verts, tets, fixed_verts), ...]: fb_create_batch — independent meshes in one context; every vector of the
        API is then the concatenation in list order."""
        self._lib = load_library()
        self._h = C.c_void_p()
        p = default_params(**params)
        self.params = p
        if batch is not None:
            vs = [_f64(m[0]).reshape(-1, 3) for m in batch]
            ts = [_i32(m[1]).reshape(-1, 4) for m in batch]
            fs = [_i32(m[2]).reshape(-1) for m in batch]
            nv, nt, nf = (_i32([len(a) for a in arrs]) for arrs in (vs, ts, fs))
            v, t = np.ascontiguousarray(np.concatenate(vs)), np.ascontiguousarray(np.concatenate(ts))
            f = np.ascontiguousarray(np.concatenate(fs)) if sum(len(a) for a in fs) else np.zeros(1, np.int32)
            st = self._lib.fb_create_batch(C.byref(self._h), len(batch), _ptr(nv), _ptr(v), _ptr(nt), _ptr(t), _ptr(nf), _ptr(f), C.byref(p))
            self._check(st, "fb_create_batch")
            self.nV, self.nT = len(v), len(t)
            self.r = self._lib.fb_num_dofs(self._h)
            return
        if veg_path is not None:
            fx = _i32(fixed_verts)
            st = self._lib.fb_create_from_veg(C.byref(self._h), str(veg_path).encode(), len(fx), _ptr(fx), C.byref(p))
            self._check(st, "fb_create_from_veg")
            self.nV, self.nT = self._lib.fb_num_vertices(self._h), self._lib.fb_num_tets(self._h)
            self.r = self._lib.fb_num_dofs(self._h)
            return
        v, t = _f64(verts).reshape(-1, 3), _i32(tets).reshape(-1, 4)
        self.nV, self.nT = len(v), len(t)
        if partition is not None:
            rank, world, comm_id = partition
            fx = _i32(fixed_verts)
            idbuf = (C.c_char * 128).from_buffer_copy(bytes(comm_id)) if comm_id is not None else None
            st = self._lib.fb_create_partitioned(C.byref(self._h), self.nV, _ptr(v), self.nT, _ptr(t), len(fx), _ptr(fx),
                                                 C.byref(p), rank, world, C.cast(idbuf, C.c_void_p) if idbuf is not None else None)
        elif constrained_dofs is not None:
            cd = _i32(constrained_dofs)
            st = self._lib.fb_create_with_constrained_dofs(C.byref(self._h), self.nV, _ptr(v), self.nT, _ptr(t), len(cd), _ptr(cd), C.byref(p))
        elif materials is not None:
            E, nu, rho = (None if m is None else _f64(m) for m in materials)
            fx = _i32(fixed_verts)
            st = self._lib.fb_create_with_materials(C.byref(self._h), self.nV, _ptr(v), self.nT, _ptr(t), len(fx), _ptr(fx),
                                                    _ptr(E) if E is not None else None, _ptr(nu) if nu is not None else None,
                                                    _ptr(rho) if rho is not None else None, C.byref(p))
        else:
            fx = _i32(fixed_verts)
            st = self._lib.fb_create(C.byref(self._h), self.nV, _ptr(v), self.nT, _ptr(t), len(fx), _ptr(fx), C.byref(p))
        self._check(st, "fb_create")
        self.r = self._lib.fb_num_dofs(self._h)

    # -- plumbing ------------------------------------------------------------------------------------
    def _check(self, st, where):
        if st != FB_OK:
            detail = self._lib.fb_last_error_string().decode() or self._lib.fb_status_string(st).decode()
            raise FemBrainError(st, where, detail)

    @property
    def batch_count(self):
        return int(self._lib.fb_batch_count(self._h))

    def batch_offsets(self):
        n = self.batch_count
        vo, to = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.int32)
        self._check(self._lib.fb_batch_offsets(self._h, _ptr(vo), _ptr(to)), "fb_batch_offsets")
        return vo, to

    def batch_cg_iterations(self):
        """Per mesh: the reference's return value (+iterations / -iterations if not converged) and rho_final / rho_0."""
        n = self.batch_count
        it, ra = np.zeros(n, np.int32), np.zeros(n)
        self._check(self._lib.fb_batch_last_cg_iterations(self._h, _ptr(it), _ptr(ra)), "fb_batch_last_cg_iterations")
        return it, ra

    @property
    def peer_memory(self):
        return bool(self._lib.fb_partition_peer_memory(self._h))

    @property
    def reordered(self):
        """1 when the row blocks were cut from a Cuthill-McKee ordering instead of the caller's numbering."""
        return int(self._lib.fb_partition_reordered(self._h))

    def partition_range(self):
        b, e = C.c_int(0), C.c_int(0)
        self._check(self._lib.fb_partition_range(self._h, C.byref(b), C.byref(e)), "fb_partition_range")
        return b.value, e.value

    def local_range(self):
        """Local vertices [lo, hi) of the inspection hooks' system that this rank owns (all of them on an ordinary context)."""
        b, e = C.c_int(0), C.c_int(0)
        self._check(self._lib.fb_partition_local_range(self._h, C.byref(b), C.byref(e)), "fb_partition_local_range")
        return b.value, e.value

    def local_to_global(self):
        """Caller's vertex id of every local vertex (identity on an ordinary context)."""
        l2g = np.zeros(max(self.local_r // 3, 1), np.int32)
        self._check(self._lib.fb_partition_local_to_global(self._h, _ptr(l2g)), "fb_partition_local_to_global")
        return l2g[: self.local_r // 3]

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.fb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    # -- sizes -----------------------------------------------------------------------------------------
    @property
    def nnz_K(self):
        return int(self._lib.fb_nnz_stiffness(self._h))

    @property
    def nnz_M(self):
        return int(self._lib.fb_nnz_mass(self._h))

    @property
    def nnz_sys(self):
        return int(self._lib.fb_nnz_system(self._h))

    @property
    def num_constrained(self):
        return self._lib.fb_num_constrained_dofs(self._h)

    @property
    def rows_sys(self):
        return self.local_r - self.num_constrained

    # -- forces / state ------------------------------------------------------------------------------------
    def set_external_forces(self, f):
        f = _f64(f).reshape(-1)
        assert f.size == self.r
        self._check(self._lib.fb_set_external_forces(self._h, _ptr(f)), "fb_set_external_forces")

    def add_external_forces(self, f):
        f = _f64(f).reshape(-1)
        assert f.size == self.r
        self._check(self._lib.fb_add_external_forces(self._h, _ptr(f)), "fb_add_external_forces")

    def set_external_forces_to_zero(self):
        self._check(self._lib.fb_set_external_forces_to_zero(self._h), "fb_set_external_forces_to_zero")

    def get_external_forces(self):
        f = np.zeros(self.r)
        self._check(self._lib.fb_get_external_forces(self._h, _ptr(f)), "fb_get_external_forces")
        return f

    def set_state(self, q, qvel=None, qaccel=None):
        q = _f64(q).reshape(-1)
        qv = _f64(qvel).reshape(-1) if qvel is not None else None
        qa = _f64(qaccel).reshape(-1) if qaccel is not None else None
        self._check(self._lib.fb_set_state(self._h, _ptr(q), _ptr(qv) if qv is not None else None,
                                           _ptr(qa) if qa is not None else None), "fb_set_state")

    def get_state(self, out=None):
        if out is None:
            out = (np.zeros(self.r), np.zeros(self.r), np.zeros(self.r))
        q, qv, qa = out
        self._check(self._lib.fb_get_state(self._h, _ptr(q) if q is not None else None, _ptr(qv) if qv is not None else None,
                                           _ptr(qa) if qa is not None else None), "fb_get_state")
        return out

    def reset_to_rest(self):
        self._check(self._lib.fb_reset_to_rest(self._h), "fb_reset_to_rest")

    def sync_force_model(self, verts, tets, fixed_verts=()):
        """Deformable::syncForceModel after a topology change: full re-setup in place, same handle, state at rest."""
        v, t, fx = _f64(verts).reshape(-1, 3), _i32(tets).reshape(-1, 4), _i32(fixed_verts)
        self._check(self._lib.fb_sync_force_model(self._h, len(v), _ptr(v), len(t), _ptr(t), len(fx), _ptr(fx)), "fb_sync_force_model")
        self.nV, self.nT = len(v), len(t)
        self.r = self._lib.fb_num_dofs(self._h)

    def set_fixed_vertices(self, fixed):
        fx = _i32(fixed)
        self._check(self._lib.fb_set_fixed_vertices(self._h, len(fx), _ptr(fx)), "fb_set_fixed_vertices")

    def set_cg(self, eps, max_iter):
        self._check(self._lib.fb_set_cg(self._h, eps, max_iter), "fb_set_cg")

    def set_grid(self, nx, ny=None, nz=None):
        """Declare the tensor grid of the mesh (CreateTruthCube numbering): what the multigrid variant coarsens."""
        ny, nz = (nx if ny is None else ny), (nx if nz is None else nz)
        self._check(self._lib.fb_set_grid(self._h, nx, ny, nz), "fb_set_grid")

    def set_solver(self, variant, warm_start=False):
        """variant: 0 / "jacobi" (the reference's solver), 1 / "block_jacobi", 2 / "mg" — labelled variants, see fembrain_b200.h."""
        code = {"jacobi": 0, "block_jacobi": 1, "mg": 2}.get(variant, variant)
        self._check(self._lib.fb_set_solver(self._h, int(code), int(bool(warm_start))), "fb_set_solver")

    def solver(self):
        v, w, lv = C.c_int(0), C.c_int(0), C.c_int(0)
        self._check(self._lib.fb_get_solver(self._h, C.byref(v), C.byref(w), C.byref(lv)), "fb_get_solver")
        nv, nb = np.zeros(max(lv.value, 1), np.int32), np.zeros(max(lv.value, 1), np.int64)
        self._check(self._lib.fb_get_solver_levels(self._h, lv.value, _ptr(nv), _ptr(nb)), "fb_get_solver_levels")
        nu, ch, ns, al = C.c_int(0), C.c_int(0), C.c_int(0), C.c_double(0)
        self._check(self._lib.fb_get_solver_smoother(self._h, C.byref(nu), C.byref(ch), C.byref(al), C.byref(ns)), "fb_get_solver_smoother")
        return {"variant": v.value, "name": self._lib.fb_solver_name(v.value).decode(), "warm_start": bool(w.value), "levels": lv.value,
                "level_vertices": [int(x) for x in nv[:lv.value]], "level_blocks": [int(x) for x in nb[:lv.value]],
                "smoothing_sweeps": nu.value, "chebyshev": bool(ch.value), "chebyshev_alpha": al.value, "structured_levels": ns.value}

    def set_warp(self, warp):
        """CorotationalLinearFEMForceModel(fem, warp): 0 linear, 1 corotational (default), 2 exact tangent."""
        self._check(self._lib.fb_set_warp(self._h, int(warp)), "fb_set_warp")

    @property
    def warp(self):
        return self._lib.fb_get_warp(self._h)

    def set_timestep(self, h):
        self._check(self._lib.fb_set_timestep(self._h, h), "fb_set_timestep")

    def set_damping(self, dm, dk):
        self._check(self._lib.fb_set_damping(self._h, dm, dk), "fb_set_damping")

    # -- the step --------------------------------------------------------------------------------------------
    def do_timestep(self):
        """VolumeConservingIntegrator::DoTimestep.  Returns 0; raises FemBrainError on solver failure."""
        self._check(self._lib.fb_step(self._h), "fb_step")
        return 0

    def step_raw(self) -> int:
        return self._lib.fb_step(self._h)

    def deformable_timestep(self):
        self._check(self._lib.fb_deformable_timestep(self._h), "fb_deformable_timestep")

    def set_gravity(self, enabled):
        self._check(self._lib.fb_deformable_set_gravity(self._h, int(enabled)), "fb_deformable_set_gravity")

    def set_floor(self, enabled, y=0.0):
        self._check(self._lib.fb_deformable_set_floor(self._h, int(enabled), y), "fb_deformable_set_floor")

    def set_haptic_forces(self, indices, forces, in_progress=True):
        idx, f = _i32(indices), _f64(forces).reshape(-1)
        assert f.size == 3 * idx.size
        self._check(self._lib.fb_deformable_set_haptic_forces(self._h, len(idx), _ptr(idx), _ptr(f), int(in_progress)),
                    "fb_deformable_set_haptic_forces")

    def set_haptic_neighborhood(self, rings):
        self._check(self._lib.fb_deformable_set_haptic_neighborhood(self._h, rings), "fb_deformable_set_haptic_neighborhood")

    def set_edge_list(self, edges, reference_quirk=True):
        e = _i32(edges).reshape(-1)
        self._check(self._lib.fb_deformable_set_edge_list(self._h, e.size // 2, _ptr(e), int(reference_quirk)), "fb_deformable_set_edge_list")

    def pick_vertices(self, box_lo, box_hi, capacity=None):
        """Deformable::pickVertices: (indices ascending, coordinates) of the vertices whose current position is in the closed box."""
        lo, hi = _f64(box_lo), _f64(box_hi)
        cap = self.nV if capacity is None else capacity
        idx, co, n = np.zeros(max(cap, 1), np.int32), np.zeros(3 * max(cap, 1)), C.c_int(0)
        self._check(self._lib.fb_deformable_pick_vertices(self._h, _ptr(lo), _ptr(hi), cap, _ptr(idx), _ptr(co), C.byref(n)), "fb_deformable_pick_vertices")
        k = min(n.value, cap)
        return idx[:k].copy(), co[:3 * k].reshape(-1, 3).copy(), n.value

    def pick_vertex(self, world_pos):
        """Deformable::pickVertex: (index, distance, position) of the vertex nearest to world_pos (lowest index on ties)."""
        w = _f64(world_pos)
        i, d, v = C.c_int(-1), C.c_double(0), np.zeros(3)
        self._check(self._lib.fb_deformable_pick_vertex(self._h, _ptr(w), C.byref(i), C.byref(d), _ptr(v)), "fb_deformable_pick_vertex")
        return i.value, d.value, v

    @property
    def contact_count(self):
        return self._lib.fb_deformable_contact_count(self._h)

    def export_positions_float4(self, rest_xyzw=None, count=None):
        """ApplyVertexDeformations: float4 rest + float4(q, 0) for the first `count` vertices."""
        n = self.nV if count is None else count
        rest = np.ascontiguousarray(rest_xyzw, dtype=np.float32).reshape(-1, 4) if rest_xyzw is not None else None
        out = np.zeros((n, 4), np.float32)
        self._check(self._lib.fb_export_positions_float4(self._h, n, _ptr(rest) if rest is not None else None, _ptr(out)),
                    "fb_export_positions_float4")
        return out

    # -- statistics ------------------------------------------------------------------------------------------
    def assembly_time(self):
        return self._lib.fb_force_assembly_seconds(self._h)

    def solve_time(self):
        return self._lib.fb_system_solve_seconds(self._h)

    def step_time(self):
        return self._lib.fb_step_seconds(self._h)

    @property
    def last_cg_iterations(self):
        return self._lib.fb_last_cg_iterations(self._h)

    @property
    def last_cg_residual_ratio(self):
        return self._lib.fb_last_cg_residual_ratio(self._h)

    @property
    def kernel_launches(self):
        return int(self._lib.fb_kernel_launches(self._h))

    @property
    def device_bytes(self):
        return int(self._lib.fb_device_bytes(self._h))

    # -- inspection ---------------------------------------------------------------------------------------------
    def K_csr(self, values=False):
        ia, ja = np.zeros(self.local_r + 1, np.int32), np.zeros(self.nnz_K, np.int32)
        self._check(self._lib.fb_get_stiffness_csr(self._h, _ptr(ia), _ptr(ja)), "fb_get_stiffness_csr")
        return ia, ja, None

    def K_row_pointers(self):
        """ia only (row pointers of K in the reference's CSR layout) — no nnz-sized column array."""
        ia = np.zeros(self.local_r + 1, np.int32)
        self._check(self._lib.fb_get_stiffness_csr(self._h, _ptr(ia), None), "fb_get_stiffness_csr")
        return ia

    @property
    def local_r(self):
        """DOFs of the LOCAL system the inspection hooks describe (= r except on a partitioned context)."""
        return int(self._lib.fb_num_local_dofs(self._h))

    def M_csr(self):
        n = self.nnz_M
        ia, ja, a = np.zeros(self.r + 1, np.int32), np.zeros(n, np.int32), np.zeros(n)
        self._check(self._lib.fb_get_mass_csr(self._h, _ptr(ia), _ptr(ja), _ptr(a)), "fb_get_mass_csr")
        return ia, ja, a

    def sys_csr(self, values=True):
        n = self.nnz_sys
        ia, ja = np.zeros(self.rows_sys + 1, np.int32), np.zeros(n, np.int32)
        a = np.zeros(n) if values else None
        self._check(self._lib.fb_get_system_csr(self._h, _ptr(ia), _ptr(ja), _ptr(a) if values else None), "fb_get_system_csr")
        return ia, ja, a

    def element_maps(self):
        row, col = np.zeros(4 * self.nT, np.int32), np.zeros(16 * self.nT, np.int32)
        self._check(self._lib.fb_get_element_maps(self._h, _ptr(row), _ptr(col)), "fb_get_element_maps")
        return row.reshape(-1, 4), col.reshape(-1, 16)

    def element_data(self):
        mi, k0 = np.zeros(16 * self.nT), np.zeros(144 * self.nT)
        self._check(self._lib.fb_get_element_data(self._h, _ptr(mi), _ptr(k0)), "fb_get_element_data")
        return mi.reshape(-1, 16), k0.reshape(-1, 144)

    def super_maps(self):
        sr, si = np.zeros(self.rows_sys, np.int32), np.zeros(self.nnz_sys, np.int32)
        self._check(self._lib.fb_get_super_maps(self._h, _ptr(sr), _ptr(si)), "fb_get_super_maps")
        return sr, si

    def submatrix_map(self):
        idx = np.zeros(self.nnz_M, np.int32)
        self._check(self._lib.fb_get_submatrix_map(self._h, _ptr(idx)), "fb_get_submatrix_map")
        return idx

    def constrained_dofs(self):
        d = np.zeros(self.num_constrained, np.int32)
        self._check(self._lib.fb_get_constrained_dofs(self._h, _ptr(d)), "fb_get_constrained_dofs")
        return d

    def force_and_matrix(self, u):
        u = _f64(u).reshape(-1)
        assert u.size == self.r
        f, a = np.zeros(self.r), np.zeros(self.nnz_K)
        self._check(self._lib.fb_compute_force_and_matrix(self._h, _ptr(u), _ptr(f), _ptr(a)), "fb_compute_force_and_matrix")
        return f, a

    def K_values(self):
        a = np.zeros(self.nnz_K)
        self._check(self._lib.fb_get_effective_stiffness_values(self._h, _ptr(a)), "fb_get_effective_stiffness_values")
        return a

    def rhs(self):
        b = np.zeros(self.rows_sys)
        self._check(self._lib.fb_get_rhs(self._h, _ptr(b)), "fb_get_rhs")
        return b

    def rhs_full(self):
        """The last step's right-hand side expanded to the LOCAL DOF numbering (zeros at constrained DOFs; InsertRows)."""
        full = np.zeros(self.local_r)
        keep = np.ones(self.local_r, bool)
        keep[self.constrained_dofs()] = False
        full[keep] = self.rhs()
        return full

    def internal_forces(self):
        f = np.zeros(self.r)
        self._check(self._lib.fb_get_internal_forces(self._h, _ptr(f)), "fb_get_internal_forces")
        return f

    def qdelta(self):
        d = np.zeros(self.r)
        self._check(self._lib.fb_get_qdelta(self._h, _ptr(d)), "fb_get_qdelta")
        return d

    def solve(self, b=None, eps=1e-6, max_iter=10000):
        x = np.zeros(self.rows_sys)
        bb = _f64(b) if b is not None else None
        it = C.c_int(0)
        self._check(self._lib.fb_solve(self._h, _ptr(bb) if bb is not None else None, _ptr(x), eps, max_iter, C.byref(it)), "fb_solve")
        return x, it.value

    def sys_spmv(self, x):
        x = _f64(x)
        y = np.zeros(self.rows_sys)
        self._check(self._lib.fb_system_multiply(self._h, _ptr(x), _ptr(y)), "fb_system_multiply")
        return y

    # -- timing --------------------------------------------------------------------------------------------------------
    def timer_start(self):
        self._check(self._lib.fb_timer_start(self._h), "fb_timer_start")

    def timer_stop(self) -> float:
        s = C.c_double(0)
        self._check(self._lib.fb_timer_stop(self._h, C.byref(s)), "fb_timer_stop")
        return s.value

    def set_profiling(self, enabled=True):
        self._check(self._lib.fb_set_profiling(self._h, int(enabled)), "fb_set_profiling")

    def spmv_profile(self):
        m, n, b = C.c_double(0), C.c_int(0), C.c_double(0)
        self._check(self._lib.fb_get_spmv_profile(self._h, C.byref(m), C.byref(n), C.byref(b)), "fb_get_spmv_profile")
        return m.value, n.value, b.value

    def set_external_forces_dev(self, dev_ptr: int):
        self._check(self._lib.fb_set_external_forces_dev(self._h, C.c_void_p(dev_ptr)), "fb_set_external_forces_dev")

    def get_state_dev(self, q_ptr=None, qvel_ptr=None, qaccel_ptr=None):
        self._check(self._lib.fb_get_state_dev(self._h, C.c_void_p(q_ptr) if q_ptr else None, C.c_void_p(qvel_ptr) if qvel_ptr else None,
                                               C.c_void_p(qaccel_ptr) if qaccel_ptr else None), "fb_get_state_dev")

    def set_external_forces_ptr(self, host_ptr: int):
        """SetExternalForces from a caller-owned host buffer (e.g. pinned memory), no numpy copy."""
        self._check(self._lib.fb_set_external_forces(self._h, C.c_void_p(host_ptr)), "fb_set_external_forces")

    def set_external_forces_owned_ptr(self, host_ptr: int):
        """fb_set_external_forces_owned from a caller-owned host buffer holding this rank's rows only."""
        self._check(self._lib.fb_set_external_forces_owned(self._h, C.c_void_p(host_ptr)), "fb_set_external_forces_owned")

    def get_state_owned_ptr(self, q_ptr=None, qvel_ptr=None, qaccel_ptr=None):
        self._check(self._lib.fb_get_state_owned(self._h, C.c_void_p(q_ptr) if q_ptr else None, C.c_void_p(qvel_ptr) if qvel_ptr else None,
                                                 C.c_void_p(qaccel_ptr) if qaccel_ptr else None), "fb_get_state_owned")

    def get_state_owned(self):
        """(q, qvel) of this rank's rows [partition_range) — the whole vectors on an ordinary context."""
        b, e = self.partition_range()
        q, qv = np.zeros(3 * (e - b)), np.zeros(3 * (e - b))
        self._check(self._lib.fb_get_state_owned(self._h, _ptr(q), _ptr(qv), None), "fb_get_state_owned")
        return q, qv

    def get_state_ptr(self, q_ptr=None, qvel_ptr=None, qaccel_ptr=None):
        self._check(self._lib.fb_get_state(self._h, C.c_void_p(q_ptr) if q_ptr else None, C.c_void_p(qvel_ptr) if qvel_ptr else None,
                                           C.c_void_p(qaccel_ptr) if qaccel_ptr else None), "fb_get_state")

    # -- micro-benchmarks ---------------------------------------------------------------------------------------------
    def bench_spmv(self, repeats=20):
        s = C.c_double(0)
        self._check(self._lib.fb_bench_spmv(self._h, repeats, C.byref(s)), "fb_bench_spmv")
        return s.value

    def bench_assembly(self, repeats=5):
        s = C.c_double(0)
        self._check(self._lib.fb_bench_assembly(self._h, repeats, C.byref(s)), "fb_bench_assembly")
        return s.value

    def bench_cg_iteration(self, repeats=50):
        s = C.c_double(0)
        self._check(self._lib.fb_bench_cg_iteration(self._h, repeats, C.byref(s)), "fb_bench_cg_iteration")
        return s.value
