"""Developer tool: times the SpMV variants selectable with FEMBRAIN_B200_SPMV on the bench workload.
   python tools/spmv_variants.py [nx]   (GPU box)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, numpy as np
sys.path.insert(0, %r)
import fembrain_b200 as fb
from bench import workload
nx = int(sys.argv[1])
v, t, fixed, f = workload(nx)
sim = fb.Simulation(v, t, fixed)
sim.set_external_forces(f)
sim.do_timestep()
sim.set_profiling(True)
sim.do_timestep()
m, n, b = sim.spmv_profile()
its = sim.last_cg_iterations
print(json.dumps({"in_step_spmv_us": m * 1e6, "in_step_gbs": b / m / 1e9, "iso_spmv_us": sim.bench_spmv(50) * 1e6,
                  "iso_iter_us": sim.bench_cg_iteration(90) * 1e6, "step_ms": sim.step_time() * 1e3, "its": its,
                  "us_per_it_in_step": sim.solve_time() * 1e6 / its}))
''' % ROOT

if __name__ == "__main__":
    nx = sys.argv[1] if len(sys.argv) > 1 else "56"
    variants = sys.argv[2:] or ["default", "rows", "l2evict", "tma", "sym"]
    for var in variants:
        env = dict(os.environ, FEMBRAIN_B200_SPMV=var)
        if var == "l2evict":  # default kernel, matrix loads with the L2 evict-first hint
            env = dict(os.environ, FEMBRAIN_B200_L2EVICT="1")
        out = subprocess.run([sys.executable, "-c", CHILD, nx], env=env, capture_output=True, text=True)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
        print(var, line, flush=True)
