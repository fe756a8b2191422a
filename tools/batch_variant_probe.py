"""Developer tool: configs[3] with the multigrid variant — 32 contexts of 196,608 tets on 32 streams — for 1 .. 32 host threads.
python tools/batch_variant_probe.py   (GPU box)"""
import sys, json; sys.path.insert(0, ".")
import bench
env = bench.Env()
for th in (1, 4, 8, 16, 32):
    out = bench.batch_variant_block(env, 32, 3, 1, threads=th)
    print(th, round(out["value"], 1), out["ms_per_batch_step"], out["cg_iterations_last_step_first_meshes"], round(out["setup_seconds"], 2), flush=True)
    env.fb.trim_memory()
