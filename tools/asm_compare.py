"""Developer tool: row-gather assembly against the two-phase path on the bench workload with a warm (R != I) state:
bit-identity of K, f, Keff and the time of each.   python tools/asm_compare.py [nx ...]   (GPU box)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, hashlib, numpy as np
sys.path.insert(0, %r)
import fembrain_b200 as fb
from bench import workload
from fembrain_b200 import meshes
nx = int(sys.argv[1])
v, t, fixed, f = workload(nx)
sim = fb.Simulation(v, t, fixed)
u = meshes.warm_displacement(v).reshape(-1)
u[sim.constrained_dofs()] = 0.0
fi, K = sim.force_and_matrix(u)
hK, hf = hashlib.sha1(K.tobytes()).hexdigest()[:16], hashlib.sha1(fi.tobytes()).hexdigest()[:16]
del K
sim.set_state(u, np.zeros_like(u))
sim.set_external_forces(f)
sim.do_timestep()
hKeff = hashlib.sha1(sim.K_values().tobytes()).hexdigest()[:16]
hrhs = hashlib.sha1(sim.rhs().tobytes()).hexdigest()[:16]
sim.set_state(u, np.zeros_like(u))
sec = sim.bench_assembly(10)
print(json.dumps({"nx": nx, "tets": len(t), "asm_ms": sec * 1e3, "mtets_s": len(t) / sec / 1e6, "K": hK, "f": hf, "Keff": hKeff,
                  "rhs": hrhs, "its": sim.last_cg_iterations, "device_GB": sim.device_bytes / 1e9}))
''' % ROOT

if __name__ == "__main__":
    for nx in (sys.argv[1:] or ["56"]):
        res = {}
        for mode in ("twophase", "gather"):
            env = dict(os.environ, FEMBRAIN_B200_ASSEMBLY=mode)
            out = subprocess.run([sys.executable, "-c", CHILD, nx], env=env, capture_output=True, text=True)
            line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-600:]
            print(mode, line, flush=True)
            try:
                res[mode] = json.loads(line)
            except Exception:
                pass
        if len(res) == 2:
            same = all(res["twophase"][k] == res["gather"][k] for k in ("K", "f", "Keff", "rhs", "its"))
            print(f"nx={nx} bit-identical: {same}  speed-up {res['twophase']['asm_ms'] / res['gather']['asm_ms']:.2f}x", flush=True)
