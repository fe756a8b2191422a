"""The LABELLED solver variants (fb_set_solver: block-Jacobi PCG, multigrid-preconditioned PCG; fb_mg.cu) against the
reference's solver and the oracle.  They are not the reference's algorithm, so the contract is: the SAME linear system
(Keff and rhs bit-identical to the parity path's, hence to the reference's), the SAME stopping rule, and after both sides
converge to a tightened residual the solution within 1e-8 relative of the oracle's (BASELINE.json north_star tolerance)."""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def _sims(dims, variant, warm=False):
    import fembrain_b200 as fb

    v, t, fixed, load = cases.cube_case(*dims)
    ref = fb.Simulation(v, t, fixed)
    var = fb.Simulation(v, t, fixed)
    if variant == "mg":
        var.set_grid(*dims)
    var.set_solver(variant, warm)
    f = cases.point_load(ref.r, load)
    for s in (ref, var):
        s.set_external_forces(f)
    return v, t, fixed, ref, var


# ell: FEMBRAIN_B200_MG_ELL read by fb_set_solver — "1" forces the structured slot-major product (k_mg_spmv_ell, by default only
# on levels of >= 200,000 vertices) on every tensor-grid level, "0" the lane-per-block product everywhere
@pytest.mark.parametrize("dims,variant,ell", [((12, 12, 12), "mg", "1"), ((9, 5, 14), "mg", "1"), ((16, 16, 16), "mg", "0"), ((6, 6, 6), "mg", "1"),
                                              ((3, 3, 3), "mg", "1"), ((9, 5, 14), "mg", "0"), ((12, 12, 12), "block_jacobi", None)])
def test_variant_solves_the_reference_system_to_the_oracle_solution(port_oracle, dims, variant, ell, monkeypatch):
    if ell is not None:
        monkeypatch.setenv("FEMBRAIN_B200_MG_ELL", ell)
    v, t, fixed, ref, var = _sims(dims, variant)
    ora = port_oracle.Oracle(v, t, fixed, kind="port")
    u = cases.perturbation(v, 0.5, 3)   # rotations != I, so the coarse levels are re-assembled at a real deformation
    u[ref.constrained_dofs()] = 0.0
    f = ref.get_external_forces()
    for s in (ref, var, ora):
        s.set_state(u, np.zeros_like(u))
        s.set_external_forces(f)
        assert s.do_timestep() == 0
    assert np.array_equal(var.K_values(), ora.K_values()), "the variant must see the reference's Keff, bit for bit"
    assert np.array_equal(var.rhs(), ora.rhs()), "... and its right-hand side"
    info = var.solver()
    assert info["variant"] == {"mg": 2, "block_jacobi": 1}[variant]
    its_ref, its_var = ref.last_cg_iterations, var.last_cg_iterations
    assert 0 < its_var < its_ref, (its_var, its_ref)
    if variant == "mg" and min(dims) >= 6:
        assert info["levels"] >= 2 and its_var <= 60, (info, its_var)   # mesh independent: ~35 on every cube size
    q, qv, _ = var.get_state()
    rq, rqv, _ = ref.get_state()
    # both stopped at eps = 1e-6 of the same measure, on different Krylov paths: two such solutions differ by up to ~cond * 1e-6
    # (the 1e-8 bar applies to the tightened solves below)
    assert cases.rel_err(q, rq) <= 1e-3 and cases.rel_err(qv, rqv) <= 1e-3
    assert var.last_cg_residual_ratio <= 1e-12
    # tightened: within 1e-8 of the oracle's tightened solution of the same system
    x, it = var.solve(eps=1e-12, max_iter=20000)
    ox, oit = ora.solve(eps=1e-12, max_iter=20000)
    assert it > 0 and oit > 0
    assert cases.rel_err(x, ox) <= 1e-8, cases.rel_err(x, ox)


def test_mg_trajectory_warm_start_and_constraint_change():
    v, t, fixed, ref, var = _sims((10, 10, 10), "mg", warm=True)
    cold = []
    for k in range(4):
        ref.do_timestep()
        var.do_timestep()
        cold.append((ref.last_cg_iterations, var.last_cg_iterations))
        q, qv, _ = var.get_state()
        rq, rqv, _ = ref.get_state()
        assert cases.rel_err(q, rq) <= 1e-3 and cases.rel_err(qv, rqv) <= 1e-3, k
        var.set_state(rq, rqv, np.zeros_like(rq))
    assert all(b < a for a, b in cold), cold
    # fewer fixed vertices: the hierarchy (constraints of every level) is rebuilt, results still match the reference solver
    half = np.asarray(fixed)[: len(fixed) // 2]
    for s in (ref, var):
        s.set_fixed_vertices(half)
        s.reset_to_rest()
        s.do_timestep()
    assert cases.rel_err(var.get_state()[0], ref.get_state()[0]) <= 1e-3
    # back to the reference solver on the same context
    var.set_solver("jacobi")
    var.reset_to_rest()
    ref.reset_to_rest()
    var.do_timestep(); ref.do_timestep()
    assert var.last_cg_iterations == ref.last_cg_iterations and np.array_equal(var.get_state()[0], ref.get_state()[0])


def test_variant_error_paths():
    import fembrain_b200 as fb
    from fembrain_b200 import api

    v, t, fixed, _ = cases.cube_case(5)
    sim = fb.Simulation(v, t, fixed)
    with pytest.raises(fb.FemBrainError) as e:
        sim.set_solver("mg")                      # no grid declared
    assert e.value.status == api.FB_ERR_INVALID_ARGUMENT
    with pytest.raises(fb.FemBrainError):
        sim.set_grid(5, 5, 4)                     # wrong node count
    ve, te, fe = cases.golden_mesh("beam3")
    other = fb.Simulation(ve, te, fe)
    with pytest.raises(fb.FemBrainError):
        other.set_grid(13, 4, 4)                  # 208 vertices, but not a tensor grid in that numbering
        other.set_solver("mg")
    other.set_solver("block_jacobi")              # any mesh
    other.set_external_forces(cases.point_load(other.r, 100))
    other.do_timestep()
    assert other.last_cg_iterations > 0
    b = fb.Simulation(batch=[(v, t, fixed), (v, t, fixed)])
    with pytest.raises(fb.FemBrainError) as e:
        b.set_solver("block_jacobi")
    assert e.value.status == api.FB_ERR_NOT_SUPPORTED
    assert sim.solver()["variant"] == 0


def test_mg_recovers_from_a_lost_smoothing_interval(monkeypatch):
    """Chebyshev smoothing is a valid preconditioner only while its interval covers the spectrum of Binv A.  With the interval
    deliberately cut to half of lambda_max the first attempt cannot converge; fb_mg_pcg_solve must notice within its 300-iteration
    cap, re-estimate lambda_max, widen the interval and deliver the same solution as the reference's solver."""
    monkeypatch.setenv("FEMBRAIN_B200_MG_CHEB_HI", "0.5")
    v, t, fixed, ref, var = _sims((10, 10, 10), "mg")
    ref.do_timestep()
    var.do_timestep()
    assert var.last_cg_iterations > 300, var.last_cg_iterations          # 300 spent + the repeated solve
    assert var.last_cg_iterations < 300 + ref.last_cg_iterations
    assert cases.rel_err(var.get_state()[0], ref.get_state()[0]) <= 1e-3
    var.do_timestep(); ref.do_timestep()
    assert 0 < var.last_cg_iterations < 100                              # the widened interval stays
