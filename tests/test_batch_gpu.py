"""A batch of independent meshes in one context (fb_create_batch, BASELINE config 4) against one context per mesh and the
oracle: everything that defines each mesh's linear system bit-exact, each mesh's PCG with its own scalars and stopping rule."""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def _meshes():
    out = []
    for spec in ((5,), (7,), (5, 3, 9), (4,)):
        v, t, fixed, load = cases.cube_case(*spec)
        out.append((v, t, fixed, load))
    return out


def test_batch_matches_one_context_per_mesh(port_oracle):
    import fembrain_b200 as fb

    ms = _meshes()
    batch = fb.Simulation(batch=[(v, t, fx) for v, t, fx, _ in ms])
    singles = [fb.Simulation(v, t, fx) for v, t, fx, _ in ms]
    vo, to = batch.batch_offsets()
    assert batch.batch_count == len(ms) and vo[-1] == sum(len(m[0]) for m in ms) and to[-1] == sum(len(m[1]) for m in ms)
    # different loads and different deformed states per mesh, so the meshes need different numbers of iterations
    f = np.concatenate([cases.point_load(3 * len(v), load, (1e4 * (k + 1), 0.0, 30.0 * k)) for k, (v, t, fx, load) in enumerate(ms)])
    u = np.concatenate([cases.perturbation(v, 0.3 * k, k) for k, (v, t, fx, load) in enumerate(ms)])
    for k, s in enumerate(singles):
        lo, hi = 3 * vo[k], 3 * vo[k + 1]
        uk = u[lo:hi].copy()
        uk[s.constrained_dofs()] = 0.0
        u[lo:hi] = uk
        s.set_state(uk, np.zeros_like(uk))
        s.set_external_forces(f[lo:hi])
    batch.set_state(u, np.zeros_like(u))
    batch.set_external_forces(f)
    for step in range(3):
        assert batch.do_timestep() == 0
        its, ratios = batch.batch_cg_iterations()
        q, qv, _ = batch.get_state()
        rhs = batch.rhs()
        fint = batch.internal_forces()
        ro = 0
        for k, s in enumerate(singles):
            lo, hi = 3 * vo[k], 3 * vo[k + 1]
            assert s.do_timestep() == 0
            srhs = s.rhs()
            # the linear system of every mesh is bit-identical to its own context's
            assert np.array_equal(fint[lo:hi], s.internal_forces()), (step, k)
            assert np.array_equal(rhs[ro:ro + len(srhs)], srhs), (step, k)
            ro += len(srhs)
            sq, sqv, _ = s.get_state()
            assert abs(int(its[k]) - s.last_cg_iterations) <= max(3, s.last_cg_iterations // 50), (its[k], s.last_cg_iterations)
            assert ratios[k] <= 1e-12
            assert cases.rel_err(q[lo:hi], sq) <= 1e-4 and cases.rel_err(qv[lo:hi], sqv) <= 1e-4
            s.set_state(q[lo:hi], qv[lo:hi])  # same trajectory for the next step's bit-exact checks
        assert len(set(int(i) for i in its)) > 1  # the meshes really stopped at different iterations
        assert batch.last_cg_iterations == int(its.max())
    # converged to the same solution: both driven to eps = 1e-12
    x, it = batch.solve(eps=1e-12, max_iter=20000)
    ro = 0
    for k, s in enumerate(singles):
        sx, sit = s.solve(eps=1e-12, max_iter=20000)
        assert cases.rel_err(x[ro:ro + len(sx)], sx) <= 1e-8
        ro += len(sx)


def test_batch_against_oracle_and_errors(port_oracle):
    import fembrain_b200 as fb
    from fembrain_b200 import api

    v, t, fx, load = cases.cube_case(6)
    batch = fb.Simulation(batch=[(v, t, fx)] * 3)
    ora = port_oracle.Oracle(v, t, fx, kind="port")
    f1 = cases.point_load(3 * len(v), load)
    batch.set_external_forces(np.concatenate([f1, 2 * f1, np.zeros_like(f1)]))  # the third mesh has a zero right-hand side
    ora.set_external_forces(f1)
    assert batch.do_timestep() == 0 and ora.do_timestep() == 0
    n = 3 * len(v)
    K = batch.K_values()
    assert np.array_equal(K[: len(K) // 3], ora.K_values())
    assert np.array_equal(batch.rhs()[: batch.rows_sys // 3], ora.rhs())
    its, _ = batch.batch_cg_iterations()
    # no load from rest: fint is rounding noise (1e-9), not 0, so the third mesh iterates on that noise like its own context
    quiet = fb.Simulation(v, t, fx)
    quiet.set_external_forces(np.zeros_like(f1))
    assert quiet.do_timestep() == 0
    assert its[0] > 0 and abs(int(its[2]) - quiet.last_cg_iterations) <= max(3, quiet.last_cg_iterations // 50)
    q = batch.get_state()[0]
    oq = ora.get_state()[0]
    assert cases.rel_err(q[:n], oq) <= 1e-4 and np.abs(q[2 * n:]).max() <= 1e-6 * np.abs(oq).max()
    # an exactly zero right-hand side: rho0 = 0, the loop condition is false at once (CGSolver.cpp:150), x stays 0
    b = batch.rhs()
    m = len(b) // 3
    b[2 * m:] = 0.0
    x, _ = batch.solve(b)
    its, _ = batch.batch_cg_iterations()
    assert its[2] == 0 and its[0] > 0 and np.all(x[2 * m:] == 0) and np.any(x[:m] != 0)
    assert cases.rel_err(q[n:2 * n], 2 * oq) <= 1e-4  # linear from rest
    with pytest.raises(fb.FemBrainError) as e:
        fb.Simulation(batch=[(v, t, fx), (v, t + 1000, fx)])
    assert e.value.status == api.FB_ERR_BAD_MESH
    # a mesh that cannot converge in the allowed iterations makes the step fail, the others are unaffected
    bad = fb.Simulation(batch=[(v, t, fx)] * 2, cg_max_iterations=5)
    bad.set_external_forces(np.concatenate([f1, np.zeros_like(f1)]))
    with pytest.raises(fb.FemBrainError) as e:
        bad.do_timestep()
    assert e.value.status == api.FB_ERR_SOLVER_NOT_CONVERGED
    its, _ = bad.batch_cg_iterations()
    assert its[0] == -5 and its[1] in (0, -5)  # the unloaded mesh iterates on rounding noise in fint; it cannot finish in 5 either


def test_step_many_equals_sequential_steps():
    """fb_step_many: independent contexts stepped from several host threads give the bits of sequential fb_step calls;
    a duplicate or NULL entry is rejected; a failing context is reported without stopping the others."""
    import ctypes as C

    import fembrain_b200 as fb
    from fembrain_b200 import api

    v, t, fixed, load = cases.cube_case(6)
    sims, refs = [], []
    for k in range(5):
        f = cases.point_load(3 * len(v), load)
        f *= 1.0 + 0.25 * k
        for group in (sims, refs):
            s = fb.Simulation(v, t, fixed)
            if k == 4:
                s.set_grid(6)
                s.set_solver("mg")
            s.set_external_forces(f)
            group.append(s)
    for _ in range(3):
        assert fb.step_many(sims, 3) == [0] * 5
        for s in refs:
            s.do_timestep()
    for a, b in zip(sims, refs):
        assert a.last_cg_iterations == b.last_cg_iterations
        assert np.array_equal(a.get_state()[0], b.get_state()[0])
    lib = api.load_library()
    dup = (C.c_void_p * 2)(sims[0]._h, sims[0]._h)
    assert lib.fb_step_many(dup, 2, 2, None) == api.FB_ERR_INVALID_ARGUMENT
    # one context that cannot converge (max 1 iteration): its status comes back, the others still step
    sims[1].set_cg(1e-6, 1)
    with pytest.raises(fb.FemBrainError) as e:
        fb.step_many(sims, 4)
    assert e.value.status == api.FB_ERR_SOLVER_NOT_CONVERGED and "context 1" in str(e.value)
