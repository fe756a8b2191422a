"""Operation-order check of fembrain_b200/csrc/fb_element_math.h without a GPU: the header is
instantiated for the host by tests/host_math_check.cpp (test-only build, -ffp-contract=off) and an
in-order assembly with it must reproduce the oracle's K and f BIT-FOR-BIT.  The product library has
no host path; on the GPU the same header is compiled with -fmad=false (tests/test_parity_gpu.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import cases

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm():
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    so = os.path.join(HERE, "_build", "libhostmath.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "host_math_check.cpp")], check=True)
    return C.CDLL(so)


def lame(E, nu):
    return (nu * E) / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))


@pytest.mark.parametrize("case", ["cube5", "two_tetra", "beam3", "eggshell"])
def test_host_instantiation_is_bit_exact(hm, port_oracle, case):
    if case == "cube5":
        v, t, fixed, _ = cases.cube_case(5)
    elif case == "two_tetra":
        from fembrain_b200 import meshes
        v, t = meshes.two_tetra()
        fixed = [0]
    else:
        v, t, fixed = cases.golden_mesh(case)
    o = port_oracle.Oracle(v, t, fixed, kind="port")
    lam, mu = lame(1e7, 0.46)
    nT, nV = len(t), len(v)
    G, K0 = np.zeros((nT, 12)), np.zeros((nT, 144))
    p = lambda a: C.c_void_p(a.ctypes.data)
    hm.hm_element_data(C.c_int(nT), p(t), p(v), C.c_double(lam), C.c_double(mu), p(G), p(K0))
    mi, k0 = o.element_data()
    assert np.array_equal(G, mi.reshape(-1, 4, 4)[:, :, :3].reshape(-1, 12))
    assert np.array_equal(K0, k0)
    ia, _, _ = o.K_csr(values=False)
    _, col = o.element_maps()
    col = np.ascontiguousarray(col)
    for seed, scale in ((1, 1.0), (2, 6.0)):
        u = cases.perturbation(v, scale, seed)
        f_ref, K_ref = o.force_and_matrix(u)
        f, Ka, its = np.zeros(3 * nV), np.zeros(o.nnz_K), np.zeros(nT, np.int32)
        hm.hm_assemble(C.c_int(nV), C.c_int(nT), p(t), p(v), p(u), C.c_double(lam), C.c_double(mu), C.c_double(1e-6), p(ia), p(col),
                       p(f), p(Ka), p(its))
        assert np.array_equal(f, f_ref) and np.array_equal(Ka, K_ref)
        assert its.min() >= 1


@pytest.mark.parametrize("case", ["cube5", "two_tetra", "eggshell"])
@pytest.mark.parametrize("warp", [0, 1, 2])
def test_host_warp_modes_bit_exact_vs_compiled_reference(hm, ref_oracle, case, warp):
    """warp = 0 (linear) / 1 / 2 (exact tangent, corotationalLinearFEM.cpp:296-428) through fbm::element_full — the function the
    device kernel k_element_full runs per element — against ComputeForceAndStiffnessMatrix(u, f, K, warp) of the compiled reference."""
    if case == "cube5":
        v, t, fixed, _ = cases.cube_case(5)
    elif case == "two_tetra":
        from fembrain_b200 import meshes
        v, t = meshes.two_tetra()
        fixed = [0]
    else:
        v, t, fixed = cases.golden_mesh(case)
    o = ref_oracle.Oracle(v, t, fixed, kind="ref")
    lam, mu = lame(1e7, 0.46)
    nT, nV = len(t), len(v)
    ia, _, _ = o.K_csr(values=False)
    _, col = o.element_maps()
    col = np.ascontiguousarray(col)
    p = lambda a: C.c_void_p(a.ctypes.data)
    for seed, scale in ((1, 1.0), (2, 6.0)):
        u = cases.perturbation(v, scale, seed)
        f_ref, K_ref = o.force_and_matrix_warp(u, warp)
        f, Ka = np.zeros(3 * nV), np.zeros(o.nnz_K)
        hm.hm_assemble_warp(C.c_int(nV), C.c_int(nT), p(t), p(v), p(u), C.c_double(lam), C.c_double(mu), C.c_double(1e-6), p(ia), p(col),
                            C.c_int(warp), p(f), p(Ka))
        assert np.array_equal(f, f_ref), np.abs(f - f_ref).max()
        assert np.array_equal(Ka, K_ref), (np.abs(Ka - K_ref).max(), np.abs(K_ref).max())


@pytest.mark.parametrize("name", ["cube5", "egg_shell_sample"])
def test_host_warp_modes_vs_committed_golden(hm, port_oracle, name):
    """The committed reference outputs for warp = 0 / 2 (tests/golden/warp_*.npz) against the host instantiation: pins the
    fixtures the GPU test uses on a box without /root/reference."""
    gold = os.path.join(HERE, "golden")
    if name == "cube5":
        v, t, fixed, _ = cases.cube_case(5)
    else:
        m = np.load(os.path.join(gold, "mesh_egg_shell_sample.npz"))
        v, t, fixed = m["verts"], m["tets"], m["fixed"]
    g = np.load(os.path.join(gold, f"warp_{name}.npz"))
    o = port_oracle.Oracle(v, t, fixed, kind="port")
    lam, mu = lame(1e7, 0.46)
    nT, nV = len(t), len(v)
    ia, _, _ = o.K_csr(values=False)
    _, col = o.element_maps()
    col = np.ascontiguousarray(col)
    p = lambda a: C.c_void_p(a.ctypes.data)
    for warp in (0, 2):
        for seed, scale in ((1, 1.0), (2, 6.0)):
            u = cases.perturbation(v, scale, seed)
            f, Ka = np.zeros(3 * nV), np.zeros(o.nnz_K)
            hm.hm_assemble_warp(C.c_int(nV), C.c_int(nT), p(t), p(v), p(u), C.c_double(lam), C.c_double(mu), C.c_double(1e-6), p(ia), p(col),
                                C.c_int(warp), p(f), p(Ka))
            assert np.array_equal(f, g[f"f_w{warp}_s{seed}"]) and np.array_equal(Ka, g[f"K_w{warp}_s{seed}"])
