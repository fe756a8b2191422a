"""Developer tool: assembly time on the 1M- and 10M-tet bench cubes, isolated (10 back-to-back assemblies) and inside a step
(step minus solve).  python tools/assembly_time.py   (GPU box)"""
import sys; sys.path.insert(0, ".")
import fembrain_b200 as fb
from bench import workload
for nx in (56, 120):
    v, t, fixed, f = workload(nx)
    sim = fb.Simulation(v, t, fixed)
    sim.set_external_forces(f)
    sim.do_timestep()
    print(nx, "assembly isolated ms", round(1e3 * sim.bench_assembly(10), 3), "step-solve ms", round(1e3 * (sim.step_time() - sim.solve_time()), 3), flush=True)
    sim.close()
