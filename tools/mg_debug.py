import sys, numpy as np
sys.path.insert(0, "/root/repo")
import fembrain_b200 as fb
from tests import cases
name, nx = sys.argv[1], int(sys.argv[2])
eps = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-6
v, t, fixed, load = cases.cube_case(nx)
sim = fb.Simulation(v, t, fixed)
if name == "mg": sim.set_grid(nx)
sim.set_solver(name)
sim.set_cg(eps, 300)
sim.set_external_forces(cases.point_load(sim.r, load))
rc = sim.step_raw()
print(name, nx, "rc", rc, "iterations", sim.last_cg_iterations, "ratio", sim.last_cg_residual_ratio, sim.solver())
