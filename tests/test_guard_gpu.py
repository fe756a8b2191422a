"""Out-of-bounds-write check of every kernel family with the library's own guard bands (FEMBRAIN_B200_GUARD=1: each device
allocation sits between two 256-byte bands of 0xA5, fb_check_guards reads them back).  Stands in for compute-sanitizer's
memcheck, which is closed on the GPU pool this was developed on (profiles/r02_sanitizer.txt); shared-memory races are covered
by the bit-reproducibility assertions (test_fullsize_gpu.py, tests/dist_check.py) instead of racecheck."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, os, sys
import numpy as np
sys.path.insert(0, %r)
import fembrain_b200 as fb
from fembrain_b200 import api
from tests import cases
report = {}
def check(tag):
    n, bad = api.check_guards()
    report[tag] = [n, bad]
# single meshes: structured, ragged slab, unstructured numbering with per-vertex valences from 4 to 40+
for name, mesh in (("cube7", cases.cube_case(7)[:3]), ("slab", cases.cube_case(5, 3, 9)[:3]), ("eggshell", cases.golden_mesh("eggshell")),
                   ("beam3", cases.golden_mesh("beam3"))):
    v, t, fixed = mesh
    sim = fb.Simulation(v, t, fixed)
    if os.environ.get("FB_GUARD_TEST_MG") == "1":   # the labelled solver variants and their level hierarchy
        if name in ("cube7", "slab"):
            sim.set_grid(*((7, 7, 7) if name == "cube7" else (5, 3, 9)))
            sim.set_solver("mg")
        else:
            sim.set_solver("block_jacobi")
    load = int(np.argmax(v[:, 1] * 1000 + v[:, 0]))
    sim.set_external_forces(cases.point_load(sim.r, load))
    for _ in range(2):
        sim.do_timestep()
    sim.force_and_matrix(cases.perturbation(v, 1.0, 1))
    sim.solve(eps=1e-10, max_iter=20000)
    sim.set_haptic_forces([load], [[1e3, 0, 0]], True)
    sim.set_floor(True, float(v[:, 1].min()) + 0.1)
    sim.deformable_timestep()
    sim.pick_vertices((-10, -10, -10), (10, 10, 10))
    sim.export_positions_float4()
    sim.set_fixed_vertices(np.asarray(fixed)[: max(1, len(fixed) // 2)])
    sim.do_timestep()
    check(name)
    sim.close()
# a batch context
v, t, fixed = cases.cube_case(5)[:3]
b = fb.Simulation(batch=[(v, t, fixed), cases.cube_case(4)[:3], (v, t, fixed)])
f = np.zeros(b.r); f[3 * (len(v) - 1)] = 1e4
b.set_external_forces(f)
b.do_timestep()
check("batch")
b.close()
print("GUARD_REPORT " + json.dumps(report))
''' % ROOT


@pytest.mark.parametrize("variant", ["default", "twophase", "mg"])
def test_no_kernel_writes_outside_its_buffers(variant):
    env = dict(os.environ, FEMBRAIN_B200_GUARD="1")
    if variant == "twophase":
        env["FEMBRAIN_B200_ASSEMBLY"] = "twophase"
    if variant == "mg":
        env["FB_GUARD_TEST_MG"] = "1"
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("GUARD_REPORT ")][-1]
    rep = json.loads(line[len("GUARD_REPORT "):])
    for tag, (n, bad) in rep.items():
        assert n >= 30, f"{tag}: guard mode not active ({n} allocations checked)"
        assert bad == 0, f"{tag}: {bad} of {n} allocations had a guard band overwritten"
