"""Developer tool: assembly kernels on the bench workload with a warm (R != I) state.  python tools/asm_profile.py [nx]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fembrain_b200 as fb  # noqa: E402
from bench import workload  # noqa: E402
from fembrain_b200 import meshes  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 56
v, t, fixed, f = workload(nx)
sim = fb.Simulation(v, t, fixed)
u = meshes.warm_displacement(v).reshape(-1)
u[sim.constrained_dofs()] = 0.0
sim.set_state(u, np.zeros_like(u))
sec = sim.bench_assembly(5)
print(f"nx={nx} tets={len(t)} assembly {sec*1e3:.3f} ms = {len(t)/sec/1e6:.1f} Mtets/s", flush=True)
