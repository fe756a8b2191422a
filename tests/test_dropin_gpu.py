"""The drop-in, run: the reference's OWN Deformable.cpp compiled with its integrator and force model replaced by the
CUDA-backed Vega subclasses (include/fembrain_b200_vega_classes.hpp; oracle/stubs/prelude_dropin.h applies the three
substitutions of INTEGRATION.md with the preprocessor) against (1) the same class compiled as the reference wrote it
(CPU) and (2) the ctypes path over the same C ABI."""
import numpy as np
import pytest

from tests import cases
from tests import test_oracle_deformable as tod

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def po(ref_oracle):
    if not (ref_oracle.dropin_available() and ref_oracle.deformable_available()):
        pytest.skip("oracle/_ref does not hold the compiled Deformable / the drop-in build")
    return ref_oracle


@pytest.mark.parametrize("name", ["cube6_low_index", "cube5_gravity_contact", "cube5_far_floor", "egg_shell_sample"])
def test_reference_deformable_on_cuda_matches_reference_deformable_on_cpu(po, name):
    v, t, fixed, hidx, hf, gravity, floor_y, rings, steps = tod.scenario(name)
    cpu, gpu = po.RefDeformable(v, t, fixed), po.RefDeformable(v, t, fixed, lib="dropin")
    for d in (cpu, gpu):
        d.set_gravity(gravity)
        d.set_floor(floor_y)
        d.set_haptic_radius(rings)
        d.set_haptic(hidx, hf, True)
    for k in range(steps):
        cpu.timestep()
        gpu.timestep()
        assert np.array_equal(cpu.external_forces(), gpu.external_forces()), f"external forces, frame {k}"
        assert cpu.contacts == gpu.contacts, f"contacts, frame {k}"
        (q, qv, qa), (gq, gqv, gqa) = cpu.get_state(), gpu.get_state()
        assert not gqa.any()
        assert cases.rel_err(gq, q) <= 1e-4 and cases.rel_err(gqv, qv) <= 1e-4, f"state, frame {k}"   # both solves stop at eps = 1e-6
        assert np.array_equal(gpu.positions(), v + gq.reshape(-1, 3))
        gpu.set_state(q, qv)   # same inputs for the next frame
    cpu.close()
    gpu.close()


def test_dropin_q_is_bit_identical_to_the_ctypes_path(po):
    """Same library underneath: the C++ class path (IntegratorBase arrays -> fb_set_state / fb_set_external_forces / fb_step /
    fb_get_state) and the ctypes path must agree bit for bit."""
    import fembrain_b200 as fb

    v, t, fixed, load = cases.cube_case(7)
    gpu = po.RefDeformable(v, t, fixed, lib="dropin")
    gpu.set_floor(-64.0)   # no contacts: the post-step leaves q alone and rewrites only v_y
    gpu.set_haptic([load], [[1e4, 0, 2e3]], True)
    sim = fb.Simulation(v, t, fixed)
    q0, qv0 = np.zeros(sim.r), np.zeros(sim.r)
    for k in range(3):
        gpu.timestep()
        sim.set_state(q0, qv0, np.zeros(sim.r))
        sim.set_external_forces(gpu.external_forces())
        sim.do_timestep()
        sq, sqv, _ = sim.get_state()
        gq, gqv, _ = gpu.get_state()
        assert np.array_equal(gq, sq), f"q, frame {k}"
        assert np.array_equal(gqv.reshape(-1, 3)[:, [0, 2]], sqv.reshape(-1, 3)[:, [0, 2]]), f"v_x, v_z, frame {k}"
        assert np.array_equal(gqv.reshape(-1, 3)[:, 1], (sqv.reshape(-1, 3)[:, 1] - sqv.reshape(-1, 3)[:, 1]) - sqv.reshape(-1, 3)[:, 1] * 0.4)
        q0, qv0 = gq, gqv
    gpu.close()


def test_force_model_subclass_matches_reference_force_model(po, port_oracle):
    """CudaCorotationalForceModel through the drop-in Deformable's own members is exercised by every step above; here
    the values it hands Vega (GetForceAndMatrix into a reference SparseMatrix) are those of the pinned oracle, bit for bit:
    the C ABI call underneath is fb_compute_force_and_matrix."""
    import fembrain_b200 as fb

    v, t, fixed, _ = cases.cube_case(5)
    sim = fb.Simulation(v, t, fixed)
    ora = port_oracle.Oracle(v, t, fixed, kind="port")
    u = cases.perturbation(v, 1.0, 2)
    f, K = sim.force_and_matrix(u)
    of, oK = ora.force_and_matrix(u)
    assert np.array_equal(f, of) and np.array_equal(K, oK)
