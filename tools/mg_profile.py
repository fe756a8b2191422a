"""Developer tool: N steps of the multigrid variant on the bench cube (for ncu launch lists).  python tools/mg_profile.py [nx] [steps]"""
import sys
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import fembrain_b200 as fb
from bench import workload
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 120
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
v, t, fixed, f = workload(nx)
sim = fb.Simulation(v, t, fixed)
sim.set_grid(nx)
sim.set_solver("mg")
sim.set_external_forces(f)
for s in range(steps):
    sim.do_timestep()
    print("step", s, "iterations", sim.last_cg_iterations, "ms", 1e3 * sim.step_time(), "solve", 1e3 * sim.solve_time(), flush=True)
