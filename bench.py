#!/usr/bin/env python
"""bench.py — the reference's headline metric on its own configurations, on B200.

Metric (BASELINE.json): implicit FEM steps/s (one VolumeConservingIntegrator::DoTimestep = corotational assembly +
Jacobi-PCG solve), with assembly Mtets/s and the SpMV's achieved HBM GB/s beside it.

Workload of the headline line (BASELINE.json configs[2], the configuration the 30 steps/s target is quoted on):
CreateTruthCube(120,120,120, 0.2) = 10,110,954 tets, E=1e7, nu=0.46, rho=1000, h=0.0333, dampK=0.01, y=0 plane fixed,
pick-mode haptic load (1e4,0,0) on the far corner node, FP64, from rest.
  N = 1 : that mesh on one GPU.
  N > 1 : THE SAME mesh split by row blocks over the N ranks (fb_create_partitioned: halo of the search direction and the
          two dot products exchanged through peer memory over NVLink, NCCL as the fallback) — strong scaling; rank 0 also
          steps the whole mesh on its own GPU and the line carries a `parity_check` block (non-zero exit if it fails).
Every line also carries, measured in the same run:
  config5_50M   CreateTruthCube(204) = 50,192,562 tets on the same N GPUs (BASELINE.json configs[4])
  config4_batch 32 independent 196,608-tet meshes per GPU in one batch context per GPU (configs[3]; N = 8: the 256 meshes)
  config2_1M    (N = 1) the 998,250-tet cube of configs[1]
  solver_variant (N = 1) the headline mesh again with the LABELLED multigrid-preconditioned CG variant (not the reference's
                algorithm; same system, same stopping rule), side by side with the parity path's iterations and ms/step
  cpu_baseline  (N = 1) the unmodified reference's DoTimestep on the host, bounded sample
`--replicas` restores round 1's N > 1 behaviour (every rank steps its own copy: weak scaling, no communication).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nx 120]

One JSON line on stdout (rank 0).  `value` = steps/s of the whole job with forces resident in HBM; `e2e` = the same
through the C ABI with HOST buffers (pinned force upload + displacement download inside the timed region every step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fem_steps_per_s"
UNIT = "steps/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def workload(nx):
    from fembrain_b200 import meshes

    v, t = meshes.truth_cube(nx)
    fixed = meshes.cube_bottom_vertices(nx)
    f = np.zeros(3 * len(v))
    f[3 * meshes.cube_corner_vertex(nx)] = 1e4
    return v, t, fixed, f


CONFIG_NAMES = {56: "BASELINE.json configs[1]", 120: "BASELINE.json configs[2]", 33: "BASELINE.json configs[3] mesh", 204: "BASELINE.json configs[4]"}


def config_dict(nx, nT, n_gpus, partitioned):
    # matrix bytes per GPU in the solver's own format (8.44 B per nonzero, ~22.5 nonzeros per tet on the cube) against the 126 MB L2
    mat_mb = 8.44 * 22.5 * (nT / (n_gpus if partitioned else 1)) / 1e6
    l2 = (f"inputs larger than L2: the 8.4 B/nnz matrix (~{mat_mb:.0f} MB per GPU) streams from HBM every SpMV; no flush needed" if mat_mb > 126
          else f"matrix ~{mat_mb:.0f} MB per GPU against a 126 MB L2 and nothing flushed between iterations: part of it can stay "
               f"L2-resident from one SpMV to the next, as it does in production use")
    return {
        "workload": f"CreateTruthCube({nx},{nx},{nx},0.2): {nT} tets, corotational FEM + Jacobi-PCG, FP64, y=0 fixed, "
                    f"point load (1e4,0,0), from rest" + (f" [{CONFIG_NAMES[nx]}]" if nx in CONFIG_NAMES else ""),
        "tets_per_gpu": nT if not partitioned else nT // n_gpus,
        "parallelism": (f"row-block partition of ONE mesh over {n_gpus} GPUs (slabs), halo of d + two dot products per PCG iteration "
                        f"exchanged through peer memory over NVLink (NCCL fallback)" if partitioned else
                        ("independent mesh per GPU, no communication (--replicas)" if n_gpus > 1 else "single GPU")),
        "l2_policy": l2,
        "cg": "eps 1e-6, max 10000, x0 = 0, exact residual every 30 iterations (reference defaults)",
    }


# ------------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (oracle/_ref) — or the pinned plain-C port when _ref is absent
def _oracle_kind():
    from oracle import pyoracle

    if pyoracle.available("ref"):
        return "ref"
    if not pyoracle.available("port"):
        pyoracle.build("port")
    return "port"


def reference_steps(nx, n_steps, budget_s, t_origin=None):
    """The reference's own setup chain and its stock VolumeConservingIntegrator::DoTimestep
    (PS_VolumeConservingIntegrator.cpp:46-260) on CreateTruthCube(nx), from rest, n_steps times (fewer if the wall-clock
    budget runs out: the number actually run is reported).  Single-threaded: the reference path has no threads
    (SURVEY.md §2b).  Assembly and solve times are the reference's own PerformanceCounters."""
    from oracle import pyoracle

    t_origin = time.perf_counter() if t_origin is None else t_origin
    kind = _oracle_kind()
    v, t, fixed, f = workload(nx)
    t0 = time.perf_counter()
    o = pyoracle.Oracle(v, t, fixed, kind=kind)
    t_setup = time.perf_counter() - t0
    log(f"[cpu] {kind} setup of nx={nx} ({len(t)} tets): {t_setup:.1f} s")
    o.set_external_forces(f)
    secs, asm, sol = [], [], []
    for s in range(n_steps):
        if secs and (time.perf_counter() - t_origin) + 1.3 * max(secs) > budget_s:
            break
        t0 = time.perf_counter()
        rc = o.do_timestep()
        secs.append(time.perf_counter() - t0)
        asm.append(o.assembly_time()); sol.append(o.solve_time())
        log(f"[cpu] step {s}: {secs[-1]:.1f} s (assembly {asm[-1]:.2f}, solve {sol[-1]:.2f}), rc {rc}")
        if rc != 0:
            break
    nnz = o.nnz_K
    o.close()
    return {"kind": "reference" if kind == "ref" else "port", "nx": nx, "tets": len(t), "nnz_K": nnz, "steps_run": len(secs),
            "seconds_per_step": float(np.mean(secs)), "step_seconds": secs, "assembly_seconds": asm, "solve_seconds": sol,
            "setup_seconds": t_setup, "assembly_mtets_per_s": len(t) / float(np.mean(asm)) / 1e6 if asm and np.mean(asm) > 0 else None}


def run_reference(args):
    """--impl reference: a RUN of the reference's stock DoTimestep on the arm's own mesh, not a model.  At 10M tets one
    reference step is several minutes, so the number of steps actually run (from rest, no warm-up) is what fits the
    wall-clock budget and is what the line reports as `steps`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t_origin = time.perf_counter()
    nx = args.nx
    nT = 6 * (nx - 1) ** 3
    want = max(1, args.steps)
    res = reference_steps(nx, want, args.ref_budget_s, t_origin)
    n = res["steps_run"]
    sec = res["seconds_per_step"]
    sample = (f"stock VolumeConservingIntegrator::DoTimestep of the compiled reference on the same mesh (nx={nx}, {nT} tets), {n} step(s) from rest, "
              f"no warm-up ({want} requested; wall-clock budget {args.ref_budget_s:.0f} s incl. {res['setup_seconds']:.0f} s of reference constructors); "
              f"1 thread of {os.cpu_count()} (the reference path is single-threaded); the reference's CG return value is not visible through "
              f"DoTimestep (mapped to 0, PS_VolumeConservingIntegrator.cpp:198-199): compare with cg_iterations_warmup[:{n}] of the GPU arm")
    line = {
        "impl": "reference", "metric": METRIC, "value": 1.0 / sec, "unit": UNIT, "n_gpus": args.gpus, "steps": n, "warmup": 0,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "weak" if (args.gpus > 1 and args.replicas) else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(nx, nT, 1, False),
        "cpu_baseline": {"value": 1.0 / sec, "unit": UNIT, "cores": 1, "kind": res["kind"], "sample": sample},
        "e2e": {"value": 1.0 / sec, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "assembly_mtets_per_s": res["assembly_mtets_per_s"], "step_seconds": res["step_seconds"],
        "assembly_seconds": res["assembly_seconds"], "solve_seconds": res["solve_seconds"],
        "setup_seconds": res["setup_seconds"], "gpu_launches": 0,
    }
    line["config"]["parallelism"] = "reference CPU path, 1 thread"
    args.emit(line)
    return 0


def cpu_baseline_sample(nx_full, nT_full, nnz_full, mean_iters_full, nx_sample):
    """cpu_baseline of the GPU arm's line: the reference's stock DoTimestep on a SMALLER cube of the same family (bounded:
    ~10-30 s of CPU work), one step from rest; `value` scales its measured assembly time by the tet ratio and its measured
    solve time by nnz ratio x iteration ratio to the bench mesh — stated in `sample`.  The full-size run of the reference is
    `bench.py --impl reference`."""
    from oracle import pyoracle

    kind = _oracle_kind()
    v, t, fixed, f = workload(nx_sample)
    t0 = time.perf_counter()
    o = pyoracle.Oracle(v, t, fixed, kind=kind)
    t_setup = time.perf_counter() - t0
    o.set_external_forces(f)
    t0 = time.perf_counter()
    o.do_timestep()
    t_step = time.perf_counter() - t0
    t_asm, t_sol = o.assembly_time(), o.solve_time()
    _, its = o.solve(eps=1e-6)   # the same system again, to read the solver's own iteration count (CGSolver.cpp:189)
    nnz_s, nT_s = o.nnz_K, len(t)
    # the reference's stronger assembly baseline, labelled separately: CorotationalLinearFEMMT, pthreads
    mt = None
    if kind == "ref" and hasattr(o, "mt_assembly_seconds"):
        try:
            threads = max(1, min(os.cpu_count() or 1, 16))
            sec_mt, fdiff = o.mt_assembly_seconds(np.zeros(3 * len(v)), threads, 2)
            mt = {"kind": "CorotationalLinearFEMMT::ComputeForceAndStiffnessMatrix (corotationalLinearFEMMT.cpp:126-177)", "threads": threads,
                  "mtets_per_s": nT_s / sec_mt / 1e6, "seconds_on_sample": sec_mt, "max_force_diff_vs_single_thread": fdiff}
        except Exception as e:  # noqa: BLE001
            mt = {"error": str(e)[:200]}
    o.close()
    est = t_asm * nT_full / nT_s + t_sol * (nnz_full / nnz_s) * (mean_iters_full / max(its, 1)) + (t_step - t_asm - t_sol) * nT_full / nT_s
    return {"value": 1.0 / est, "unit": UNIT, "cores": 1, "kind": "reference" if kind == "ref" else "port",
            "sample": (f"stock DoTimestep of the compiled reference on CreateTruthCube({nx_sample}) = {nT_s} tets, one step from rest: {t_step:.2f} s "
                       f"(assembly {t_asm:.2f} s, solve {t_sol:.2f} s, {its} PCG iterations; constructors {t_setup:.1f} s); scaled to the bench mesh "
                       f"({nT_full} tets, {mean_iters_full:.0f} iterations/step) by tets (assembly) and nnz x iterations (solve); 1 thread of "
                       f"{os.cpu_count()}; the full-size reference run is `bench.py --impl reference`"),
            "measured_steps_per_s_on_sample": 1.0 / t_step, "assembly_mtets_per_s": nT_s / t_asm / 1e6,
            "seconds_per_cg_iteration_on_sample": t_sol / max(its, 1), "sample_tets": nT_s, "assembly_multithreaded": mt}


def ncu_table(key, field):
    """Figures from the committed `ncu --set full` captures (profiles/ncu_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            return json.load(fh).get(str(key), {}).get(field)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------
class Env:
    """torch / torch.distributed plumbing shared by the GPU legs."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        import fembrain_b200 as fb

        self.torch, self.dist, self.fb = torch, dist, fb
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: the CUDA library has no fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return vals
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return tuple(float(x) for x in t)

    def sum_over_ranks(self, val):
        if self.world == 1:
            return int(val)
        t = self.torch.tensor([int(val)], dtype=self.torch.int64, device="cuda")
        self.dist.all_reduce(t)
        return int(t[0])

    def comm_id(self):
        idt = self.torch.zeros(128, dtype=self.torch.uint8, device="cuda")
        if self.rank == 0:
            idt.copy_(self.torch.frombuffer(bytearray(self.fb.comm_unique_id()), dtype=self.torch.uint8))
        self.dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


class Stepper:
    """K steps of one context, resident or end to end, timed on the device between barriers."""

    def __init__(self, env, sim, f, partitioned, tets=None):
        torch = env.torch
        self.env, self.sim, self.partitioned = env, sim, partitioned
        self.r = sim.r
        if partitioned:
            b, e = sim.partition_range()
            if sim.reordered:
                order, _ = env.fb.partition_ordering(sim.nV, tets, env.world)   # (the cube keeps its numbering: slabs)
                own = f.reshape(-1, 3)[order[b:e]].reshape(-1)
            else:
                own = f[3 * b:3 * e]
            self.f_pinned = torch.from_numpy(np.ascontiguousarray(own)).pin_memory()
            self.q_pinned = torch.empty(3 * (e - b), dtype=torch.float64).pin_memory()
            sim.set_external_forces(f)           # resident arm: the force vector stays on every rank
            self.f_dev = None
        else:
            self.f_pinned = torch.from_numpy(f).pin_memory()
            self.q_pinned = torch.empty(self.r, dtype=torch.float64).pin_memory()
            self.f_dev = torch.from_numpy(f).cuda()
        self.h2d = self.f_pinned.numel() * 8
        self.d2h = self.q_pinned.numel() * 8
        torch.cuda.synchronize()

    def region(self, n_steps, e2e):
        sim = self.sim
        iters, t_asm, t_solve = [], 0.0, 0.0
        self.env.barrier()
        sim.timer_start()
        for _ in range(n_steps):
            if e2e:   # H2D of this step's forces from pinned host memory
                if self.partitioned:
                    sim.set_external_forces_owned_ptr(self.f_pinned.data_ptr())
                else:
                    sim.set_external_forces_ptr(self.f_pinned.data_ptr())
            elif not self.partitioned:
                sim.set_external_forces_dev(self.f_dev.data_ptr())   # resident in HBM
            sim.do_timestep()
            if e2e:   # D2H of the displacement vector
                if self.partitioned:
                    sim.get_state_owned_ptr(self.q_pinned.data_ptr())
                else:
                    sim.get_state_ptr(self.q_pinned.data_ptr())
            iters.append(int(sim.last_cg_iterations))
            t_asm += sim.assembly_time(); t_solve += sim.solve_time()
        sec = sim.timer_stop()
        self.env.barrier()
        return sec, iters, t_asm, t_solve

    def owned_q(self):
        return self.sim.get_state_owned()[0]


def make_sim(env, v, t, fixed, partitioned, comm_id=None):
    fb = env.fb
    if partitioned:
        return fb.Simulation(v, t, fixed, partition=(env.rank, env.world, comm_id), device=env.local)
    return fb.Simulation(v, t, fixed, device=env.local)


def measure_mesh(env, nx, steps, warmup, partitioned, full=True, want_parity=False):
    """The whole measurement of one cube on the job's GPUs.  Returns (line-dict on rank 0 or None, extras)."""
    world, rank = env.world, env.rank
    t0 = time.perf_counter()
    v, t, fixed, f = workload(nx)
    t_mesh = time.perf_counter() - t0
    nT, r = len(t), 3 * len(v)
    comm_id = env.comm_id() if partitioned else None
    env.barrier()
    t0 = time.perf_counter()
    sim = make_sim(env, v, t, fixed, partitioned, comm_id)
    env.barrier()
    t_setup = time.perf_counter() - t0
    st = Stepper(env, sim, f, partitioned, t)

    sim.reset_to_rest()
    _, it_warm, _, _ = st.region(warmup, False)
    sim.set_profiling(True)
    l0 = sim.kernel_launches
    with ClockSampler(env.local) as clk:
        sec, iters, t_asm, t_solve = st.region(steps, False)
    launches = sim.kernel_launches - l0
    spmv_mean, spmv_samples, spmv_bytes = sim.spmv_profile()
    sim.set_profiling(False)
    q_end_owned = st.owned_q()
    traj = None
    if want_parity and partitioned:
        from tests import dist_parity
        traj = (it_warm + iters, dist_parity.global_state(sim, r)[0])

    e2e = None
    if full:
        sim.reset_to_rest()
        st.region(warmup, True)
        sec_e2e, iters_e2e, _, _ = st.region(steps, True)
        same_work = iters_e2e == iters
        same_state = bool(np.array_equal(st.q_pinned.numpy(), q_end_owned))
        if world == 1:
            assert same_work, "e2e arm must do the same work as the resident arm"
            assert same_state, "e2e arm must end in the same state"
        e2e = (sec_e2e, same_work, same_state)

    iso = None
    if full:
        # isolated kernel timings on the final matrices (back to back, matrix >> L2)
        t_spmv_iso = sim.bench_spmv(50)
        t_iter_iso = sim.bench_cg_iteration(60) if not partitioned else float("nan")
        t_asm_iso = sim.bench_assembly(5)
        iso = (t_spmv_iso, t_iter_iso, t_asm_iso)

    sec, = env.max_over_ranks(sec)
    if e2e:
        e2e = (env.max_over_ranks(e2e[0])[0],) + e2e[1:]
    launches = env.sum_over_ranks(launches)
    dev_bytes = sim.device_bytes
    nnz_local, peer = sim.nnz_K, (bool(sim.peer_memory) if partitioned else None)
    nnz = env.sum_over_ranks(nnz_local) if partitioned else nnz_local  # (cut rows are counted on both sides: an upper bound)

    out = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        rows_local = sim.local_r
        units = 1 if (partitioned or world == 1) else world
        mean_it = float(np.mean(iters))
        achieved = spmv_bytes / spmv_mean / 1e9 if spmv_mean > 0 else None
        b_iter = spmv_bytes + 72.0 * rows_local   # one PCG iteration moves B_spmv + 72 B/row beyond the SpMV, own-format bytes
        refresh = sum(i // 30 for i in iters)
        solve_bytes = sum(iters) * b_iter + refresh * spmv_bytes
        out = {
            "value": units * steps / sec, "ms_per_step": 1e3 * sec / steps,
            "cg_iterations_per_step": iters, "cg_iterations_warmup": it_warm, "mean_iterations": mean_it,
            "ms_per_cg_iteration": 1e3 * sec / max(sum(iters), 1),
            "assembly_mtets_per_s": units * nT * steps / t_asm / 1e6 if t_asm > 0 else None,
            "assembly_share_of_step": t_asm / sec, "solve_share_of_step": t_solve / sec,
            "setup_seconds": t_setup, "mesh_generation_seconds": t_mesh, "device_bytes_rank0": dev_bytes, "nnz_K": nnz, "tets": nT,
            "gpu_launches": launches, "clocks": clk.summary(), "peer_memory": peer,
            "roofline": {
                "kernel": "k_spmv_rows3<1> (q = Keff d fused with d.q), CUDA-event pairs around every 16th PCG iteration's launch inside the timed steps"
                          + (" (rank 0's rows; includes the wait for the neighbours' halo)" if partitioned else ""),
                "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": ncu_table(nx, "spmv_bytes_per_launch") if not partitioned else None,
                "algorithmic_bytes_per_launch": spmv_bytes, "mean_launch_seconds": spmv_mean, "samples": spmv_samples,
                "bytes_model": "8 B/nnz values + 4 B per 3x3 block column + 52 B per block row (rowptr, x read, y write)",
                "reference_layout_equiv_gbs": ((12.0 * nnz_local + 20.0 * rows_local) / spmv_mean / 1e9) if spmv_mean > 0 else None,
                "pcg_iteration_gbs_in_step": solve_bytes / t_solve / 1e9 if t_solve > 0 else None,
                "pcg_iteration_frac_in_step": solve_bytes / t_solve / 1e9 / peak if t_solve > 0 else None,
            },
        }
        if e2e:
            out["e2e"] = {"value": units * steps / e2e[0], "unit": UNIT, "h2d_bytes_per_step": env.world * st.h2d if partitioned else st.h2d * units,
                          "d2h_bytes_per_step": env.world * st.d2h if partitioned else st.d2h * units, "ms_per_step": 1e3 * e2e[0] / steps,
                          "same_iterations_as_resident": e2e[1], "same_final_state_as_resident": e2e[2],
                          "api": ("fb_set_external_forces_owned / fb_step / fb_get_state_owned per rank (each rank moves its own rows)" if partitioned
                                  else "fb_set_external_forces / fb_step / fb_get_state")}
        if iso:
            out["roofline"]["isolated_spmv_gbs"] = spmv_bytes / iso[0] / 1e9
            if iso[1] == iso[1]:
                out["roofline"]["pcg_iteration_gbs_isolated"] = b_iter / iso[1] / 1e9
            comp = 16.0 * nT + 72.0 * (r / 3) + 8.0 * nnz
            # compulsory bytes of one assembly = tet ids + x0, u read, f write + K values written; the kernels are
            # bound by the FP64 pipe (no-FMA arithmetic, bit-identical to the reference), not by these bytes
            share = world if partitioned else 1   # a partitioned rank assembles its own rows: ~1/world of the mesh per GPU
            out["assembly"] = {"compulsory_bytes": comp, "compulsory_gbs_per_gpu": comp / share / iso[2] / 1e9,
                               "seconds_isolated": iso[2], "mtets_per_s_isolated": nT / iso[2] / 1e6,
                               "ncu": ncu_table(nx, "assembly")}
    return out, sim, traj, (v, t, fixed, f)


# ------------------------------------------------------------------------------------------------------
def batch_block(env, per_gpu, steps, warmup, nx=33):
    """BASELINE.json configs[3]: independent meshes, `per_gpu` of them per GPU in ONE batch context per GPU
    (fb_create_batch: block-diagonal system, PCG scalars and stopping rule per mesh), no communication."""
    torch, fb = env.torch, env.fb
    v, t, fixed, _ = workload(nx)
    nT, r = len(t), 3 * len(v)
    total = per_gpu * env.world
    corner = 3 * (len(v) - 1)
    forces = []
    for k in range(per_gpu):   # same mesh, load direction rotated about y with the mesh index
        m = env.rank * per_gpu + k
        ang = 2.0 * np.pi * m / max(total, 1)
        f = np.zeros(r)
        f[corner], f[corner + 2] = 1e4 * np.cos(ang), 1e4 * np.sin(ang)
        forces.append(f)
    t0 = time.perf_counter()
    sim = fb.Simulation(batch=[(v, t, fixed)] * per_gpu, device=env.local)
    t_setup = time.perf_counter() - t0
    f_all = np.concatenate(forces)
    f_dev = torch.from_numpy(f_all).cuda()
    f_pin = torch.from_numpy(f_all).pin_memory()
    q_pin = torch.empty_like(f_pin).pin_memory()

    def region(n, e2e):
        env.barrier()
        sim.timer_start()
        for _ in range(n):
            if e2e:
                sim.set_external_forces_ptr(f_pin.data_ptr())
            else:
                sim.set_external_forces_dev(f_dev.data_ptr())
            sim.do_timestep()
            if e2e:
                sim.get_state_ptr(q_pin.data_ptr())
        sec = sim.timer_stop()
        env.barrier()
        return env.max_over_ranks(sec)[0]

    region(warmup, False)
    sim.set_profiling(True)
    sec = region(steps, False)
    m_s, n_s, b_s = sim.spmv_profile()
    sim.set_profiling(False)
    its = [int(x) for x in sim.batch_cg_iterations()[0][:4]]
    sim.reset_to_rest()
    region(warmup, True)
    sec_e2e = region(steps, True)
    sim.close()
    if env.rank != 0:
        return None
    peak, _ = measured_peak_gbs()
    return {"workload": f"{total} independent meshes of CreateTruthCube({nx}) = {nT} tets each, {per_gpu} per GPU in one batch context per GPU, "
                        f"load direction rotated per mesh, no communication" + (" [BASELINE.json configs[3]: 256 meshes over 8 GPUs]" if total == 256 else ""),
            "metric": "fem_mesh_steps_per_s", "value": total * steps / sec, "unit": "mesh-steps/s", "steps": steps, "warmup": warmup,
            "ms_per_batch_step": 1e3 * sec / steps, "mtets_steps_per_s": total * steps / sec * nT / 1e6,
            "e2e": {"value": total * steps / sec_e2e, "unit": "mesh-steps/s", "h2d_bytes_per_step": 8 * r * total, "d2h_bytes_per_step": 8 * r * total},
            "cg_iterations_last_step_first_meshes": its, "setup_seconds": t_setup,
            "roofline": {"kernel": "kb_spmv<1> (all meshes of the context that still iterate), event pairs in step, rank 0", "bound": "hbm",
                         "achieved": b_s / m_s / 1e9 if m_s > 0 else None, "peak": peak, "unit": "GB/s",
                         "frac": b_s / m_s / 1e9 / peak if m_s > 0 else None, "samples": n_s, "traffic": ncu_table(f"batch{per_gpu}x{nx}", "spmv_bytes_per_launch")}}


def batch_variant_block(env, per_gpu, steps, warmup, nx=33, threads=8):
    """configs[3] with the labelled multigrid variant: the same `per_gpu` independent meshes per GPU, each its own context with
    FB_SOLVER_MG_PCG on its own stream (the batch context keeps the reference's solver), stepped by fb_step_many's pool of host
    threads so that the latency-bound cycles of different meshes overlap on the GPU.  Wall clock between two device synchronisations (the work
    runs on `per_gpu` streams), max over ranks."""
    torch, fb = env.torch, env.fb
    v, t, fixed, _ = workload(nx)
    nT, r = len(t), 3 * len(v)
    total = per_gpu * env.world
    corner = 3 * (len(v) - 1)
    sims = []
    t0 = time.perf_counter()
    for k in range(per_gpu):
        m = env.rank * per_gpu + k
        ang = 2.0 * np.pi * m / max(total, 1)
        f = np.zeros(r)
        f[corner], f[corner + 2] = 1e4 * np.cos(ang), 1e4 * np.sin(ang)
        sim = fb.Simulation(v, t, fixed, device=env.local)
        sim.set_grid(nx)
        sim.set_solver("mg")
        sim.set_external_forces(f)
        sims.append(sim)
    t_setup = time.perf_counter() - t0
    threads = max(1, min(threads, per_gpu))

    def region(n):
        env.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fb.step_many(sims, threads)   # fb_step_many: one step of every context from `threads` native host threads
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        env.barrier()
        return env.max_over_ranks(sec)[0]

    region(warmup)
    sec = region(steps)
    its = [int(s.last_cg_iterations) for s in sims[:4]]
    name = sims[0].solver()["name"]
    for sim in sims:
        sim.close()
    if env.rank != 0:
        return None
    return {"solver": name, "value": total * steps / sec, "unit": "mesh-steps/s", "steps": steps, "warmup": warmup, "host_threads_per_gpu": threads,
            "contexts_per_gpu": per_gpu, "ms_per_batch_step": 1e3 * sec / steps, "cg_iterations_last_step_first_meshes": its,
            "setup_seconds": t_setup, "timing": "wall clock between device synchronisations (one stream per mesh), max over ranks"}


def variant_block(env, nx, steps, warmup, parity_iters, parity_ms, q_parity_end):
    """The labelled solver variant on the headline mesh (single GPU): FB_SOLVER_MG_PCG — the same Keff and rhs, the same
    stopping rule, a multigrid-preconditioned CG instead of the reference's Jacobi-PCG (fembrain_b200/csrc/fb_mg.cu).  Same
    K steps from rest, resident and end to end; the final displacement is compared with the parity path's (both stop at
    eps = 1e-6 of the same weighted residual, on different Krylov paths)."""
    fb = env.fb
    v, t, fixed, f = workload(nx)
    nT, r = len(t), 3 * len(v)
    t0 = time.perf_counter()
    sim = fb.Simulation(v, t, fixed, device=env.local)
    sim.set_grid(nx)
    sim.set_solver("mg")
    t_setup = time.perf_counter() - t0
    st = Stepper(env, sim, f, False, t)
    sim.reset_to_rest()
    _, it_warm, _, _ = st.region(warmup, False)
    sec, iters, t_asm, t_solve = st.region(steps, False)
    q_end = st.owned_q()
    sim.reset_to_rest()
    st.region(warmup, True)
    sec_e2e, iters_e2e, _, _ = st.region(steps, True)
    info = sim.solver()
    nb64 = sim.nnz_K // 9
    # algorithmic bytes of ONE iteration: the FP64 product + update + direction on the finest level, and per level of the cycle
    # 2 nu FP16 products (24 B per stored block — slot-major levels store nSlots = 15 per vertex — or 28 B per block with its
    # column index; per vertex: x 16, b 16, out 16, Binv 36 in the smoothing products), the product-free first sweep
    # (68 B/vertex), restriction / prolongation (48 B per fine vertex + 32 B per coarse vertex)
    nu = max(1, info["smoothing_sweeps"])
    b_outer = 8.0 * 9 * nb64 + 4.0 * nb64 + 52.0 * (r / 3) + (61.0 + 21.0) * r
    b_cycle = 0.0
    lv, lb = info["level_vertices"], info["level_blocks"]
    for k in range(len(lv) - 1):
        b_mat = 24.0 * 15 * lv[k] if k < info["structured_levels"] else 28.0 * lb[k]
        b_cycle += 2 * nu * (b_mat + 52.0 * lv[k]) + 36.0 * lv[k] + 68.0 * lv[k] + 48.0 * lv[k] + 32.0 * lv[k + 1]
    b_iter = b_outer + b_cycle
    peak, _ = measured_peak_gbs()
    sim.close()
    out = {
        "solver": info["name"], "levels": info["levels"], "level_vertices": lv,
        "smoother": {"sweeps": nu, "chebyshev": info["chebyshev"], "alpha": info["chebyshev_alpha"], "structured_levels": info["structured_levels"]},
        "value": steps / sec, "unit": UNIT, "ms_per_step": 1e3 * sec / steps, "steps": steps, "warmup": warmup,
        "e2e": {"value": steps / sec_e2e, "unit": UNIT, "h2d_bytes_per_step": st.h2d, "d2h_bytes_per_step": st.d2h,
                "same_iterations_as_resident": iters_e2e == iters},
        "cg_iterations_per_step": iters, "cg_iterations_warmup": it_warm,
        "parity_path_iterations_per_step": parity_iters, "parity_path_ms_per_step": parity_ms,
        "speedup_vs_parity_path": parity_ms / (1e3 * sec / steps),
        "ms_per_iteration": 1e3 * t_solve / max(sum(iters), 1), "solve_share_of_step": t_solve / sec, "assembly_share_of_step": t_asm / sec,
        "setup_seconds": t_setup,
        "roofline_iteration": {"bound": "hbm", "algorithmic_bytes_per_iteration": b_iter, "achieved": b_iter * sum(iters) / t_solve / 1e9,
                               "peak": peak, "unit": "GB/s", "frac": b_iter * sum(iters) / t_solve / 1e9 / peak,
                               "note": "whole solve time (cycle set-up included) over iterations x modelled bytes; phases per iteration in profiles/r02_mg_iteration_phases.txt"},
        "final_displacement_rel_diff_vs_parity_path": float(np.abs(q_end - q_parity_end).max() / np.abs(q_parity_end).max()),
        "contract": "same Keff and rhs (bit-identical assembly), same stopping rule sum r^2/diag <= eps^2 sum b^2/diag with eps = 1e-6; "
                    "tests/test_solver_variants_gpu.py: <= 1e-8 of the oracle after both converge to eps = 1e-12",
    }
    return out


def run_ours(args):
    env = Env()
    world, rank = env.world, env.rank
    partitioned = world > 1 and not args.replicas
    nx = args.nx
    blocks_ok = True
    main, sim, traj, mesh = measure_mesh(env, nx, args.steps, args.warmup, partitioned, full=True,
                                         want_parity=partitioned and not args.no_parity)
    v, t, fixed, f = mesh
    nT, r = len(t), 3 * len(v)
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak" if (world > 1 and not partitioned) else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(nx, nT, world, partitioned),
            "solver": "jacobi_pcg (the reference's algorithm; parity path)",
        }
        for k in ("clocks", "e2e", "gpu_launches", "roofline", "cg_iterations_per_step", "cg_iterations_warmup", "ms_per_cg_iteration",
                  "assembly_mtets_per_s", "assembly_share_of_step", "solve_share_of_step", "assembly", "setup_seconds",
                  "mesh_generation_seconds", "device_bytes_rank0", "nnz_K", "tets", "peer_memory"):
            if k in main:
                line[k] = main[k]
        line["target"] = {"steps_per_s_at_10M_tets_1gpu": 30.0, "note": "north_star target; Jacobi-PCG needs ~950 iterations x 2.5 GB per step on this mesh (>= 0.36 s at the HBM roofline)"}

    # ---- multi-GPU parity inside the driver-run command --------------------------------------------------------------
    if partitioned and not args.no_parity:
        from tests import dist_parity
        ref = None
        try:
            if rank == 0:
                ref = make_sim(env, v, t, fixed, False)
                ref.set_external_forces(f)
            res = dist_parity.partition_parity(sim, ref, rank, world, r, traj_part=traj, log=log, tight_eps=1e-12 if nT < 2_000_000 else 1e-10)
            blocks_ok &= bool(res["ok"])
            if rank == 0:
                res["mesh"] = f"the bench mesh itself (nx={nx}, {nT} tets); trajectory = the {args.warmup}+{args.steps} resident steps of this run"
                line["parity_check"] = res
        finally:
            if ref is not None:
                ref.close()
    q_parity_end = sim.get_state_owned()[0] if world == 1 else None
    sim.close()
    env.fb.trim_memory()

    def guarded(name, fn):
        """Auxiliary blocks never take the headline down: an exception becomes {"error": ...} (all ranks must agree to go on)."""
        try:
            out = fn()
            ok = 1
        except Exception as e:  # noqa: BLE001
            out, ok = {"error": f"{type(e).__name__}: {e}"[:400]}, 0
        if world > 1:
            tt = env.torch.tensor([ok], device="cuda")
            env.dist.all_reduce(tt, op=env.dist.ReduceOp.MIN)
            ok = int(tt[0])
        if rank == 0 and out is not None:
            line[name] = out
        env.fb.trim_memory()
        return ok

    # ---- the labelled faster-converging solver variant on the headline mesh (single GPU) -----------------------------------
    if world == 1 and not args.no_variant:
        guarded("solver_variant", lambda: variant_block(env, nx, args.steps, args.warmup, main["cg_iterations_per_step"], main["ms_per_step"],
                                                        q_parity_end))

    # ---- configs[4]: 50M tets on the same GPUs -------------------------------------------------------------------------
    if not args.no_50m:
        def block50():
            m, s50, _, mesh50 = measure_mesh(env, args.nx50, args.steps50, args.warmup50, world > 1, full=False)
            var50 = None
            if world == 1 and not args.no_variant:
                # the labelled multigrid-preconditioned variant on the same context: same steps from rest
                try:
                    t0 = time.perf_counter()
                    s50.set_grid(args.nx50)
                    s50.set_solver("mg")
                    t_set = time.perf_counter() - t0
                    st50 = Stepper(env, s50, mesh50[3], False, mesh50[1])
                    s50.reset_to_rest()
                    _, itw, _, _ = st50.region(args.warmup50, False)
                    sec50, it50, _, tsol50 = st50.region(args.steps50, False)
                    var50 = {"solver": s50.solver()["name"], "value": args.steps50 / sec50, "unit": UNIT, "ms_per_step": 1e3 * sec50 / args.steps50,
                             "cg_iterations_per_step": it50, "cg_iterations_warmup": itw, "ms_per_iteration": 1e3 * tsol50 / max(sum(it50), 1),
                             "set_solver_seconds": t_set, "device_bytes": s50.device_bytes}
                except Exception as e:  # noqa: BLE001
                    var50 = {"error": f"{type(e).__name__}: {e}"[:300]}
            s50.close()
            if m is None:
                return None
            keep = ("value", "ms_per_step", "cg_iterations_per_step", "cg_iterations_warmup", "ms_per_cg_iteration", "assembly_mtets_per_s",
                    "setup_seconds", "mesh_generation_seconds", "device_bytes_rank0", "tets", "peer_memory", "gpu_launches", "roofline")
            out = {k: m[k] for k in keep if k in m}
            out.update({"workload": config_dict(args.nx50, m["tets"], world, world > 1)["workload"], "unit": UNIT, "steps": args.steps50,
                        "warmup": args.warmup50, "parallelism": config_dict(args.nx50, m["tets"], world, world > 1)["parallelism"]})
            if var50 is not None:
                out["solver_variant"] = var50
            return out
        guarded("config5_50M", block50)

    # ---- configs[3]: batch of independent meshes, 32 per GPU -------------------------------------------------------------
    if not args.no_batch:
        guarded("config4_batch", lambda: batch_block(env, args.batch_per_gpu, args.steps_batch, args.warmup_batch))
        if not args.no_variant:
            def batch_var():
                out = batch_variant_block(env, args.batch_per_gpu, args.steps_batch, args.warmup_batch)
                if rank == 0 and isinstance(line.get("config4_batch"), dict):
                    line["config4_batch"]["solver_variant"] = out
                return None
            guarded("config4_batch_variant", batch_var)

    # ---- configs[1] and the CPU baseline, single GPU only ------------------------------------------------------------------
    if world == 1 and not args.no_1m and nx != 56:
        def block1m():
            m, s1, _, mesh1 = measure_mesh(env, 56, 10, 3, False, full=True)
            var1 = None
            if not args.no_variant:
                try:   # the labelled variant on the same context, same 3 + 10 steps from rest
                    s1.set_grid(56)
                    s1.set_solver("mg")
                    st1 = Stepper(env, s1, mesh1[3], False, mesh1[1])
                    s1.reset_to_rest()
                    st1.region(3, False)
                    sec1, it1, _, _ = st1.region(10, False)
                    var1 = {"solver": s1.solver()["name"], "value": 10 / sec1, "unit": UNIT, "ms_per_step": 1e3 * sec1 / 10, "cg_iterations_per_step": it1}
                except Exception as e:  # noqa: BLE001
                    var1 = {"error": f"{type(e).__name__}: {e}"[:300]}
            s1.close()
            keep = ("value", "ms_per_step", "e2e", "cg_iterations_per_step", "ms_per_cg_iteration", "assembly_mtets_per_s", "setup_seconds", "tets",
                    "roofline", "assembly")
            out = {k: m[k] for k in keep if k in m}
            out.update({"workload": config_dict(56, m["tets"], 1, False)["workload"], "unit": UNIT, "steps": 10, "warmup": 3})
            if var1 is not None:
                out["solver_variant"] = var1
            return out
        guarded("config2_1M", block1m)
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_sample(nx, nT, main["nnz_K"], float(np.mean(main["cg_iterations_per_step"])), args.cpu_nx)
        except Exception as e:  # noqa: BLE001 — the checker is optional plumbing for the bench
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"unavailable: {e}"}

    if rank == 0:
        args.emit(line)
    env.close()
    return 0 if blocks_ok else 3


# ------------------------------------------------------------------------------------------------------
def run_batch(args):
    """--batch B: BASELINE.json configs[3] as the headline line (B independent meshes spread over the GPUs)."""
    env = Env()
    per = max(1, args.batch // env.world)
    with ClockSampler(env.local) as clk:
        blk = batch_block(env, per, args.steps, args.warmup, nx=args.nx if args.nx != 120 else 33)
    if env.rank == 0:
        line = {"metric": blk["metric"], "value": blk["value"], "unit": blk["unit"], "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": blk["ms_per_batch_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": blk["workload"], "parallelism": "one batch context per GPU, no communication"},
                "clocks": clk.summary(), "e2e": blk["e2e"], "roofline": blk["roofline"], "mtets_steps_per_s": blk["mtets_steps_per_s"],
                "cg_iterations_last_step_first_meshes": blk["cg_iterations_last_step_first_meshes"]}
        args.emit(line)
    env.close()
    return 0


def main():
    # stdout carries exactly ONE JSON line: everything libraries print there meanwhile (e.g. NCCL's version banner)
    # is sent to stderr, and the saved descriptor is restored for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        return _main(saved_stdout)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)


def emit(saved_stdout, line):
    sys.stdout.flush()
    os.write(saved_stdout, (json.dumps(line) + "\n").encode())


def _main(saved_stdout):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=120, help="cube resolution of the headline line: 120 = configs[2] (10.1M tets), 56 = configs[1] (1M tets)")
    ap.add_argument("--replicas", action="store_true", help="N>1: every rank steps its own copy of the mesh (weak scaling, no communication) instead of ONE partitioned mesh")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the built-in parity check of the partitioned path")
    ap.add_argument("--no-50m", action="store_true")
    ap.add_argument("--nx50", type=int, default=204, help="cube resolution of the config5_50M block (204 = 50,192,562 tets)")
    ap.add_argument("--steps50", type=int, default=3)
    ap.add_argument("--warmup50", type=int, default=1)
    ap.add_argument("--no-batch", action="store_true")
    ap.add_argument("--batch-per-gpu", type=int, default=32)
    ap.add_argument("--steps-batch", type=int, default=3)
    ap.add_argument("--warmup-batch", type=int, default=1)
    ap.add_argument("--no-1m", action="store_true")
    ap.add_argument("--no-variant", action="store_true", help="skip the solver_variant block (multigrid-preconditioned CG on the headline mesh)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-nx", type=int, default=40, help="cube resolution of the bounded cpu_baseline sample (40 = 355,914 tets)")
    ap.add_argument("--ref-budget-s", type=float, default=780.0, help="--impl reference: wall-clock budget for the whole run")
    ap.add_argument("--batch", type=int, default=0, help="headline = configs[3]: a batch of this many independent meshes (nx 33)")
    ap.add_argument("--quick", action="store_true", help="headline mesh only (no 50M / batch / 1M / cpu blocks)")
    args = ap.parse_args()
    if args.quick:
        args.no_50m = args.no_batch = args.no_1m = args.no_cpu_baseline = args.no_variant = True
    if args.warmup < 3 and args.impl == "ours":
        print("note: timing rules ask for >= 3 warm-up steps", file=sys.stderr)
    args.emit = lambda line: emit(saved_stdout, line)
    if args.impl == "reference":
        return run_reference(args)
    return run_batch(args) if args.batch > 0 else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
