"""Developer tool: fb_create wall time (second creation in the process, so module load is excluded).  python tools/setup_time.py [nx]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fembrain_b200 as fb  # noqa: E402
from bench import workload  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 56
v, t, fixed, f = workload(nx)
fb.Simulation(v, t, fixed).close()
for k in range(2):
    t0 = time.perf_counter()
    sim = fb.Simulation(v, t, fixed)
    dt = time.perf_counter() - t0
    print(f"nx={nx} tets={len(t)} fb_create {dt*1e3:.1f} ms  device {sim.device_bytes/1e9:.3f} GB", flush=True)
    sim.close()
