"""Developer tool: parity path vs solver variants on the bench cube.  python tools/variant_probe.py [nx] [steps]   (GPU box)"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import fembrain_b200 as fb  # noqa: E402
from bench import workload  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 56
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
v, t, fixed, f = workload(nx)
out = {}
ALL = (("jacobi", False), ("block_jacobi", False), ("mg", False), ("mg", True))
want = sys.argv[3].split(",") if len(sys.argv) > 3 else None
for name, warm in ALL:
    if want and (name + ("+warm" if warm else "")) not in want:
        continue
    sim = fb.Simulation(v, t, fixed)
    t0 = time.perf_counter()
    if name == "mg":
        sim.set_grid(nx)
    sim.set_solver(name, warm)
    t_set = time.perf_counter() - t0
    sim.set_external_forces(f)
    its, ms, msol = [], [], []
    for s in range(steps):
        sim.do_timestep()
        its.append(sim.last_cg_iterations); ms.append(1e3 * sim.step_time()); msol.append(1e3 * sim.solve_time())
    q = sim.get_state()[0]
    key = name + ("+warm" if warm else "")
    out[key] = {"iterations": its, "step_ms": [round(x, 2) for x in ms], "solve_ms": [round(x, 2) for x in msol], "set_solver_s": round(t_set, 3),
                "info": sim.solver(), "bytes": sim.device_bytes}
    if "jacobi" in out and key != "jacobi":
        out[key]["rel_err_vs_jacobi"] = float(np.abs(q - qj).max() / np.abs(qj).max())
    if key == "jacobi":
        qj = q
    print(key, json.dumps(out[key]), flush=True)
    sim.close()
