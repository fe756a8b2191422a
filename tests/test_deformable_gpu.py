"""GPU parity of the caller around the step — Deformable::timestep (DEF/Deformable.cpp:318-420): gravity,
haptic force ring spreading (both neighbour rules), floor-plane post-step, vertex picks.

Checkers: (1) frames of the reference's OWN compiled `Deformable` (oracle/deformable_harness.cpp; committed as
tests/golden/deformable_*.npz by make_golden.py, and run live where oracle/_ref holds it); (2) the restatement in
oracle/deformable_port.inc / pyoracle.pick_*, itself pinned bit-for-bit against (1) by tests/test_oracle_deformable.py."""
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def unique_edges(tets, seed=0):
    """A deterministic stand-in for VolMesh::m_vEdges: unique undirected edges in first-seen order."""
    seen, out = set(), []
    for t in tets:
        for a, b in ((1, 2), (2, 3), (3, 1), (2, 0), (0, 3), (0, 1)):  # maskTetEdges, VolMesh.cpp:465
            k = (min(t[a], t[b]), max(t[a], t[b]))
            if k not in seen:
                seen.add(k)
                out.append((int(t[a]), int(t[b])))
    return np.array(out, np.int32)


@pytest.mark.parametrize("name", ["cube6_low_index", "cube5_gravity_contact", "cube5_far_floor", "egg_shell_sample"])
def test_deformable_timestep_matches_compiled_reference_frames(name):
    """fb_deformable_timestep against frames produced by the reference's compiled Deformable::timestep: external forces
    (gravity, haptic rings through the get_node_neighbors quirk on VolMesh's own edge array) bit-exact, contact counts
    equal, state within the eps = 1e-6 solver tolerance; every frame restarts from the reference's state."""
    import fembrain_b200 as fb
    from tests import test_oracle_deformable as tod

    v, t, fixed, hidx, hf, gravity, _, rings, steps = tod.scenario(name)
    z = np.load(os.path.join(cases.GOLDEN, f"deformable_{name}.npz"))
    sim = fb.Simulation(v, t, fixed)
    sim.set_gravity(gravity)
    sim.set_floor(True, float(z["floor_y"]))
    sim.set_haptic_neighborhood(rings)
    sim.set_haptic_forces(np.array(hidx, np.int32), np.array(hf, np.float64), True)
    sim.set_edge_list(z["edges"], True)
    zero = np.zeros(sim.r)
    for k in range(steps):
        sim.deformable_timestep()
        assert np.array_equal(sim.get_external_forces(), z["ext"][k]), f"external forces, frame {k}"
        assert sim.contact_count == int(z["contacts"][k]), f"contacts, frame {k}"
        q, qv, qa = sim.get_state()
        assert not qa.any()
        assert cases.rel_err(q, z["q"][k]) <= 1e-4 and cases.rel_err(qv, z["qvel"][k]) <= 1e-4, f"state, frame {k}"
        sim.set_state(z["q"][k], z["qvel"][k], zero)


def test_picks_match_compiled_reference_live(ref_oracle):
    """Deformable::pickVertices / pickVertex evaluated by the compiled reference itself on ITS deformed mesh vs the device
    queries on the same displacement (skipped where oracle/_ref does not hold the compiled Deformable)."""
    import fembrain_b200 as fb

    if not ref_oracle.deformable_available():
        pytest.skip("compiled Deformable not available")
    v, t, fixed, load = cases.cube_case(6)
    d = ref_oracle.RefDeformable(v, t, fixed)
    d.set_haptic([load], [[1e4, 2e3, 0]], True)
    d.timestep()
    q, qv, _ = d.get_state()
    pos = d.positions()
    sim = fb.Simulation(v, t, fixed)
    sim.set_state(q, qv, np.zeros_like(q))
    for lo, hi in [((-0.25, 0.15, -0.25), (0.25, 0.65, 0.25)), ((-10, -10, -10), (10, 10, 10)), ((5, 5, 5), (6, 6, 6)),
                   (tuple(pos[10]), tuple(pos[10])), (tuple(pos.min(axis=0)), tuple(pos.max(axis=0)))]:
        idx, co, n = sim.pick_vertices(lo, hi)
        ridx, rco = d.pick_vertices(lo, hi)
        assert n == len(ridx) and np.array_equal(idx, ridx) and np.array_equal(co, rco)
    rng = np.random.default_rng(11)
    for w in list(rng.uniform(-1, 2, size=(8, 3))) + [pos[17], 0.5 * (pos[0] + pos[1])]:
        i, dist, p = sim.pick_vertex(w)
        ri, rd, rp = d.pick_vertex(w)
        assert i == ri and dist == rd and np.array_equal(p, rp)
    d.close()


@pytest.mark.parametrize("quirk", [False, True])
@pytest.mark.parametrize("gravity", [False, True])
def test_deformable_timestep_matches_restatement(port_oracle, quirk, gravity):
    import fembrain_b200 as fb

    v, t, fixed, load = cases.cube_case(6)
    sim = fb.Simulation(v, t, fixed)
    ora = port_oracle.Oracle(v, t, fixed, kind="port")
    edges = unique_edges(t)
    hidx = np.array([load, load - 7, 40], np.int32)
    hf = np.array([[1e4, 0, 0], [0, -3e3, 2e3], [5e2, 5e2, 5e2]], dtype=np.float64)
    floor_y = 0.35  # a plane inside the cube so that some nodes are in contact
    sim.set_gravity(gravity)
    sim.set_floor(True, floor_y)
    sim.set_haptic_neighborhood(5)
    sim.set_haptic_forces(hidx, hf, True)
    sim.set_edge_list(edges if quirk else np.zeros((0, 2), np.int32), quirk)
    contacts = 0
    for step in range(3):
        rc, contacts = ora.deformable_timestep(gravity=gravity, haptic_idx=hidx, haptic_forces=hf, in_progress=True, rings=5,
                                               edges=edges if quirk else None, quirk=quirk, floor=True, floor_y=floor_y,
                                               contacts=contacts)
        assert rc == 0
        sim.deformable_timestep()
        # the external force vector is pure host set arithmetic: bit-exact
        assert np.array_equal(sim.get_external_forces(), ora.get_external_forces()), "external forces"
        assert sim.contact_count == contacts
        q, qv, qa = sim.get_state()
        oq, oqv, oqa = ora.get_state()
        assert np.all(qa == 0) and np.all(oqa == 0)
        assert cases.rel_err(q, oq) <= 1e-4 and cases.rel_err(qv, oqv) <= 1e-4  # both solves stop at eps = 1e-6
        sim.set_state(oq, oqv, oqa)  # keep trajectories on identical inputs


def test_haptic_not_in_progress_and_rings_one(port_oracle):
    import fembrain_b200 as fb

    v, t, fixed, load = cases.cube_case(5)
    sim = fb.Simulation(v, t, fixed)
    sim.set_floor(False)
    sim.set_haptic_forces([load], [[1e4, 0, 0]], in_progress=False)  # m_bHapticInProgress == false => no forces
    sim.deformable_timestep()
    assert not sim.get_external_forces().any()
    sim.set_haptic_forces([load], [[1e4, 0, 0]], in_progress=True)
    sim.set_haptic_neighborhood(1)  # no spreading loop at all (j from 1 to < 1)
    sim.deformable_timestep()
    f = sim.get_external_forces()
    assert f[3 * load] == 1e4 and np.count_nonzero(f) == 1


def test_floor_poststep_rewrites_every_velocity(port_oracle):
    """The reference damps the normal velocity of EVERY node each frame (v_y <- -0.4 v_y), contact or not (:373-392)."""
    import fembrain_b200 as fb

    v, t, fixed, load = cases.cube_case(4)
    sim = fb.Simulation(v, t, fixed)
    ora = port_oracle.Oracle(v, t, fixed, kind="port")
    rng = np.random.default_rng(5)
    v0 = rng.standard_normal(sim.r) * 0.1
    v0[sim.constrained_dofs()] = 0
    for s in (sim, ora):
        s.set_state(np.zeros(sim.r), v0)
    sim.set_floor(True, -100.0)
    sim.deformable_timestep()
    rc, ct = ora.deformable_timestep(floor=True, floor_y=-100.0)
    assert ct == 0 and sim.contact_count == 0
    q, qv, _ = sim.get_state()
    oq, oqv, _ = ora.get_state()
    # random (high-frequency) initial velocities: two eps = 1e-6 solves that stop an iteration apart differ by
    # ~cond * 1e-6 in the solution, so the bound here is looser than in the smooth cases
    assert cases.rel_err(qv, oqv) <= 5e-3 and cases.rel_err(q, oq) <= 5e-3


def test_pick_vertices_and_pick_vertex_match_restatement(port_oracle):
    """Force producers' queries (Deformable::pickVertices / pickVertex) on the deformed positions, incl. box faces
    (closed interval), ties (lowest index), empty results and a too-small output buffer."""
    import fembrain_b200 as fb

    v, t, fixed, load = cases.cube_case(7)
    sim = fb.Simulation(v, t, fixed)
    sim.set_external_forces(cases.point_load(sim.r, load))
    sim.do_timestep()
    q = sim.get_state()[0]
    pos = v + q.reshape(-1, 3)
    boxes = [((-0.25, 0.15, -0.25), (0.25, 0.65, 0.25)), ((-10, -10, -10), (10, 10, 10)), ((5, 5, 5), (6, 6, 6)),
             (tuple(pos[10]), tuple(pos[10])),  # a degenerate box exactly on a vertex: closed interval keeps it
             (tuple(pos.min(axis=0)), tuple(pos.max(axis=0)))]
    for lo, hi in boxes:
        idx, co, n = sim.pick_vertices(lo, hi)
        oidx, oco = port_oracle.pick_vertices(v, q, lo, hi)
        assert n == len(oidx) and np.array_equal(idx, oidx) and np.array_equal(co, oco)
    idx, co, n = sim.pick_vertices((-10, -10, -10), (10, 10, 10), capacity=5)
    assert n == sim.nV and np.array_equal(idx, np.arange(5))
    rng = np.random.default_rng(3)
    for w in list(rng.uniform(-1, 2, size=(8, 3))) + [pos[17], 0.5 * (pos[0] + pos[1])]:
        i, d, p = sim.pick_vertex(w)
        oi, od, op = port_oracle.pick_vertex(v, q, w)
        assert i == oi and d == od and np.array_equal(p, op)
    # tie: a point equidistant from vertices 0 and 1 of the REST mesh -> the lower index
    sim.reset_to_rest()
    i, d, p = sim.pick_vertex(0.5 * (v[0] + v[1]))
    assert i == port_oracle.pick_vertex(v, np.zeros_like(q), 0.5 * (v[0] + v[1]))[0]
