#!/usr/bin/env python
"""bench.py — the reference's headline metric on its own configuration, on B200.

Metric (BASELINE.json): implicit FEM steps/s (one VolumeConservingIntegrator::DoTimestep = corotational
assembly + Jacobi-PCG solve), with assembly Mtets/s and the SpMV's achieved HBM GB/s beside it.
Workload at N=1 (BASELINE.json configs[1]): CreateTruthCube(56,56,56, 0.2) = 998,250 tets, E=1e7, nu=0.46,
rho=1000, h=0.0333, dampK=0.01, y=0 plane fixed, pick-mode haptic load (1e4,0,0) on the far corner node,
FP64, starting from rest.  At N>1 every rank steps its own copy of that mesh (config-4 style
batch of independent meshes: no data-path collective, weak scaling); `--partitioned` instead splits ONE mesh
by row blocks across the ranks with NCCL halo exchange (config 5, strong scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nx 56]

One JSON line on stdout (rank 0).  `value` = steps/s of the whole job with forces resident in HBM;
`e2e` = the same through the C ABI with HOST buffers (pinned force upload + displacement download
inside the timed region every step).
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


def config_dict(nx, nT, n_gpus, partitioned):
    # matrix bytes per GPU in the solver's own format (8.44 B per nonzero, ~22.5 nonzeros per tet on the cube) against the 126 MB L2
    mat_mb = 8.44 * 22.5 * (nT / (n_gpus if partitioned else 1)) / 1e6
    l2 = (f"inputs larger than L2: the 8.4 B/nnz matrix (~{mat_mb:.0f} MB per GPU) streams from HBM every SpMV; no flush needed" if mat_mb > 126
          else f"matrix ~{mat_mb:.0f} MB per GPU against a 126 MB L2 and nothing flushed between iterations: part of it can stay "
               f"L2-resident from one SpMV to the next, as it does in production use")
    return {
        "workload": f"CreateTruthCube({nx},{nx},{nx},0.2): {nT} tets, corotational FEM + Jacobi-PCG, FP64, y=0 fixed, "
                    f"point load (1e4,0,0), from rest" + (" [BASELINE.json configs[1]]" if nx == 56 else ""),
        "tets_per_gpu": nT if not partitioned else nT // n_gpus,
        "parallelism": ("row-block partition + NCCL halo" if partitioned else
                        ("independent mesh per GPU, no communication" if n_gpus > 1 else "single GPU")),
        "l2_policy": l2,
        "cg": "eps 1e-6, max 10000, x0 = 0, exact residual every 30 iterations (reference defaults)",
    }


# ------------------------------------------------------------------------------------------------------
def cpu_reference_sample(nx, cg_sample_iters, iters_per_step, steps, warmup, log=lambda *a: None):
    """Times the reference's own CPU code (oracle/_ref) — or the plain-C port when _ref is absent — on the
    same workload, bounded: per step the FULL assembly (GetForceAndMatrix) plus `cg_sample_iters` PCG
    iterations on the constrained matrix, extrapolated to `iters_per_step` iterations (the PCG cost is
    linear in the iteration count).  Single-threaded: the reference path has no threads."""
    from oracle import pyoracle

    kind = "ref" if pyoracle.available("ref") else "port"
    if kind == "port" and not pyoracle.available("port"):
        pyoracle.build("port")
    v, t, fixed, f = workload(nx)
    t0 = time.perf_counter()
    o = pyoracle.Oracle(v, t, fixed, kind=kind)
    t_setup = time.perf_counter() - t0
    log(f"[cpu] {kind} setup {t_setup:.1f} s")
    u = np.zeros(o.r)
    b = f[np.setdiff1d(np.arange(o.r), np.concatenate([3 * fixed, 3 * fixed + 1, 3 * fixed + 2]))]
    per_step = []
    t_asm_l, t_it_l = [], []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        o.force_and_matrix(u)
        t_asm = time.perf_counter() - t0
        o.load_system_from_K()
        t0 = time.perf_counter()
        o.solve_iters(cg_sample_iters, b)
        t_cg = time.perf_counter() - t0
        t_iter = t_cg / cg_sample_iters
        if s >= warmup:
            per_step.append(t_asm + iters_per_step * t_iter)
            t_asm_l.append(t_asm); t_it_l.append(t_iter)
    sec = float(np.mean(per_step))
    return {
        "kind": "reference" if kind == "ref" else "port",
        "cores": 1,
        "seconds_per_step": sec,
        "value": 1.0 / sec,
        "assembly_mtets_per_s": len(t) / float(np.mean(t_asm_l)) / 1e6,
        "seconds_per_cg_iteration": float(np.mean(t_it_l)),
        "setup_seconds": t_setup,
        "sample": (f"same mesh (nx={nx}, {len(t)} tets); per step: full GetForceAndMatrix + {cg_sample_iters} PCG iterations on the "
                   f"constrained matrix, extrapolated to {iters_per_step} iterations/step; 1 thread of {os.cpu_count()} "
                   f"(the reference path is single-threaded)"),
    }


# committed iteration counts of the bench workload (from rest, steps 1..): measured on B200 by this
# bench (matches the CPU oracle to +-1 where the oracle is affordable); used by --impl reference to
# extrapolate its bounded PCG sample without touching the GPU arm
def ncu_assembly(nx):
    """ncu figures of the assembly kernels for this size (profiles/ncu_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            return json.load(fh).get(str(nx), {}).get("assembly")
    except Exception:
        return None


def ncu_traffic(nx):
    """dram__bytes_read.sum + dram__bytes_write.sum per SpMV launch from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json), or None when no capture exists for this size."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            return json.load(fh).get(str(nx), {}).get("spmv_bytes_per_launch")
    except Exception:
        return None


def known_iterations(nx):
    try:
        with open(os.path.join(ROOT, "profiles", "bench_iterations.json")) as fh:
            tab = json.load(fh)
        return tab.get(str(nx))
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    its = known_iterations(args.nx)
    iters = int(np.mean(its[args.warmup:args.warmup + args.steps])) if its else 650
    from oracle import pyoracle  # noqa: F401
    res = cpu_reference_sample(args.nx, args.cpu_cg_iters, iters, args.steps, args.warmup, log=lambda *a: print(*a, file=sys.stderr))
    from fembrain_b200 import meshes
    nT = 6 * (args.nx - 1) ** 3
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args.nx, nT, 1, False),
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "assembly_mtets_per_s": res["assembly_mtets_per_s"], "cg_iterations_per_step_assumed": iters,
        "seconds_per_cg_iteration": res["seconds_per_cg_iteration"], "gpu_launches": 0,
    }
    args.emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import fembrain_b200 as fb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA library has no fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nx = args.nx
    v, t, fixed, f = workload(nx)
    nT, r = len(t), 3 * len(v)
    t0 = time.perf_counter()
    partitioned = args.partitioned and world > 1
    if partitioned:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(fb.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        sim = fb.Simulation(v, t, fixed, partition=(rank, world, bytes(idt.cpu().numpy().tobytes())), device=local)
    else:
        sim = fb.Simulation(v, t, fixed, device=local)
    t_setup = time.perf_counter() - t0

    f_pinned = torch.from_numpy(f).pin_memory()
    q_pinned = torch.empty(r, dtype=torch.float64).pin_memory()
    f_dev = torch.from_numpy(f).cuda()
    torch.cuda.synchronize()

    def timed_region(n_steps, e2e):
        iters, t_asm, t_solve = [], 0.0, 0.0
        barrier()
        sim.timer_start()
        for _ in range(n_steps):
            if e2e:
                sim.set_external_forces_ptr(f_pinned.data_ptr())      # H2D from pinned host memory
            elif not partitioned:
                sim.set_external_forces_dev(f_dev.data_ptr())         # resident in HBM
            # (partitioned: the force vector set before the region stays resident on every rank)
            sim.do_timestep()
            if e2e:
                sim.get_state_ptr(q_pinned.data_ptr())                # D2H of the displacement vector
            iters.append(sim.last_cg_iterations)
            t_asm += sim.assembly_time(); t_solve += sim.solve_time()
        sec = sim.timer_stop()
        barrier()
        return sec, iters, t_asm, t_solve

    # warm-up (from rest), then K timed steps resident, then state reset and the same K steps end to end
    sim.reset_to_rest()
    if partitioned:
        sim.set_external_forces(f)
    _, it_warm, _, _ = timed_region(args.warmup, False)
    sim.set_profiling(True)
    l0 = sim.kernel_launches
    with ClockSampler(local) as clk:
        sec, iters, t_asm, t_solve = timed_region(args.steps, False)
    launches = sim.kernel_launches - l0
    spmv_mean, spmv_samples, spmv_bytes = sim.spmv_profile()
    sim.set_profiling(False)
    q_end = sim.get_state()[0]

    sim.reset_to_rest()
    timed_region(args.warmup, True)
    sec_e2e, iters_e2e, _, _ = timed_region(args.steps, True)
    assert iters_e2e == iters, "e2e arm must do the same work as the resident arm"
    assert np.array_equal(q_pinned.numpy(), q_end), "e2e arm must end in the same state"

    # isolated kernel timings on the final matrices (back to back, matrix >> L2)
    t_spmv_iso = sim.bench_spmv(50)
    t_iter_iso = sim.bench_cg_iteration(60) if not partitioned else float("nan")
    t_asm_iso = sim.bench_assembly(5)

    # max over ranks
    if world > 1:
        tt = torch.tensor([sec, sec_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec, sec_e2e = float(tt[0]), float(tt[1])
        ll = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ll)
        launches = int(ll[0])

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        nnz = sim.nnz_K
        rows = r
        units = 1 if partitioned else world  # partitioned: one mesh; otherwise independent meshes stepped concurrently
        value = units * args.steps / sec
        e2e_value = units * args.steps / sec_e2e
        achieved = spmv_bytes / spmv_mean / 1e9 if spmv_mean > 0 else None
        mean_it = float(np.mean(iters))
        # one PCG iteration moves B_spmv + 72 B/row beyond the SpMV, own-format bytes
        b_iter = spmv_bytes + 72.0 * rows
        refresh = sum(i // 30 for i in iters)
        solve_bytes = sum(iters) * b_iter + refresh * spmv_bytes
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "strong" if partitioned else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(nx, nT, world, partitioned),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * r, "d2h_bytes_per_step": 8 * r,
                    "ms_per_step": 1e3 * sec_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": {
                "kernel": "k_spmv_rows3<1> (q = Keff d fused with d.q), CUDA-event pairs around every 16th PCG iteration's launch inside the timed steps",
                "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic(nx),
                "algorithmic_bytes_per_launch": spmv_bytes, "mean_launch_seconds": spmv_mean, "samples": spmv_samples,
                "bytes_model": "8 B/nnz values + 4 B per 3x3 block column + 52 B per block row (rowptr, x read, y write)",
                "reference_layout_equiv_gbs": ((12.0 * nnz + 20.0 * rows) / spmv_mean / 1e9) if spmv_mean > 0 else None,
                "isolated_spmv_gbs": spmv_bytes / t_spmv_iso / 1e9,
                "pcg_iteration_gbs_in_step": solve_bytes / t_solve / 1e9 if t_solve > 0 else None,
                "pcg_iteration_gbs_isolated": b_iter / t_iter_iso / 1e9,
            },
            "cg_iterations_per_step": iters, "cg_iterations_warmup": it_warm,
            "assembly_mtets_per_s": units * nT * args.steps / t_asm / 1e6 if t_asm > 0 else None,
            "assembly_mtets_per_s_isolated": nT / t_asm_iso / 1e6,
            "assembly_share_of_step": t_asm / sec, "solve_share_of_step": t_solve / sec,
            # compulsory bytes of one assembly = tet ids + x0, u read, f write + K values written; the kernels are
            # bound by the FP64 pipe (no-FMA arithmetic, bit-identical to the reference), not by these bytes
            "assembly": {"compulsory_bytes": 16.0 * nT + 72.0 * (rows / 3) + 8.0 * nnz,
                         "compulsory_gbs": (16.0 * nT + 72.0 * (rows / 3) + 8.0 * nnz) / t_asm_iso / 1e9,
                         "seconds_isolated": t_asm_iso, "ncu": ncu_assembly(nx)},
            "setup_seconds": t_setup, "device_bytes": sim.device_bytes, "nnz_K": nnz, "tets": nT,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_reference_sample(nx, args.cpu_cg_iters, int(round(mean_it)), 1, 0)
                line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"],
                                        "assembly_mtets_per_s": cb["assembly_mtets_per_s"],
                                        "seconds_per_cg_iteration": cb["seconds_per_cg_iteration"]}
            except Exception as e:  # the checker is optional plumbing for the bench
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"unavailable: {e}"}
        args.emit(line)
    sim.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------------
def run_batch(args):
    """BASELINE.json configs[3]: a batch of independent meshes spread over the GPUs, no communication.  Every rank owns
    batch/world contexts (each with its own CUDA stream) and drives them from `--streams` host threads, so the launch gaps
    of one small mesh are filled by the kernels of another."""
    import torch
    import torch.distributed as dist

    import fembrain_b200 as fb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA library has no fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx = args.nx
    v, t, fixed, f0 = workload(nx)
    nT, r = len(t), 3 * len(v)
    mine = [m for m in range(args.batch) if m % world == rank]
    if args.group < 0:
        args.group = max(1, len(mine))
    sims, forces = [], []
    corner = 3 * (len(v) - 1)
    for m in mine:  # same mesh, load direction rotated about y with the mesh index
        ang = 2.0 * np.pi * m / max(args.batch, 1)
        f = np.zeros(r)
        f[corner], f[corner + 2] = 1e4 * np.cos(ang), 1e4 * np.sin(ang)
        if args.group > 0:
            forces.append(f)
        else:
            sims.append(fb.Simulation(v, t, fixed, device=local))
            forces.append(torch.from_numpy(f).cuda())
    if args.group > 0:
        # fb_create_batch: `group` meshes per context as one block-diagonal system, PCG scalars and stopping rule per mesh
        grouped = []
        for g0 in range(0, len(mine), args.group):
            n = min(args.group, len(mine) - g0)
            sims.append(fb.Simulation(batch=[(v, t, fixed)] * n, device=local))
            grouped.append(torch.from_numpy(np.concatenate(forces[g0:g0 + n])).cuda())
        forces = grouped
    # end-to-end arm: the same forces from pinned host memory every step, every mesh's displacements read back
    f_pinned = [x.cpu().pin_memory() for x in forces]
    q_pinned = [torch.empty_like(x) for x in f_pinned]
    q_pinned = [x.pin_memory() for x in q_pinned]
    torch.cuda.synchronize()
    nthreads = max(1, min(args.streams, len(sims)))
    iters = [[] for _ in sims]

    def worker(tid, nsteps, record, e2e):
        for _ in range(nsteps):
            for k in range(tid, len(sims), nthreads):
                if e2e:
                    sims[k].set_external_forces_ptr(f_pinned[k].data_ptr())
                else:
                    sims[k].set_external_forces_dev(forces[k].data_ptr())
                sims[k].do_timestep()
                if e2e:
                    sims[k].get_state_ptr(q_pinned[k].data_ptr())
                if record:
                    iters[k].append(int(sims[k].batch_cg_iterations()[0][0]) if args.group > 0 else sims[k].last_cg_iterations)

    def region(nsteps, record, e2e=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=worker, args=(i, nsteps, record, e2e)) for i in range(nthreads)]
        [x.start() for x in th]
        [x.join() for x in th]
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sec = float(tt[0])
        return sec

    region(args.warmup, False)
    l0 = sum(s_.kernel_launches for s_ in sims)
    for s_ in sims:
        s_.set_profiling(True)
    with ClockSampler(local) as clk:
        sec = region(args.steps, True)
    launches = sum(s_.kernel_launches for s_ in sims) - l0
    prof = [s_.spmv_profile() for s_ in sims]
    for s_ in sims:
        s_.set_profiling(False)
    # the same steps end to end: state back to rest, warm-up, K timed steps with host buffers
    for s_ in sims:
        s_.reset_to_rest()
    region(args.warmup, False, True)
    sec_e2e = region(args.steps, False, True)
    if world > 1:
        ll = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ll)
        launches = int(ll[0])
    if rank == 0:
        value = args.batch * args.steps / sec
        cfg = config_dict(nx, nT, world, False)
        cfg["workload"] = f"batch of {args.batch} independent meshes, each " + cfg["workload"] + " (load direction rotated per mesh)"
        cfg["parallelism"] = f"{args.batch // world} meshes per GPU on {nthreads} host threads/streams, no communication"
        if args.group > 0:
            cfg["parallelism"] = (f"{args.batch // world} meshes per GPU in batch contexts of {args.group} (fb_create_batch: block-diagonal "
                                  f"system, PCG scalars and stopping rule per mesh), {nthreads} host threads/streams, no communication")
        cfg["timing"] = "host wall clock between device synchronisations (many streams), max over ranks"
        h2d = sum(x.numel() * 8 for x in f_pinned)
        line = {
            "metric": "fem_mesh_steps_per_s", "value": value, "unit": "mesh-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clk.summary(),
            "e2e": {"value": args.batch * args.steps / sec_e2e, "unit": "mesh-steps/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": h2d * world, "ms_per_step": 1e3 * sec_e2e / args.steps},
            "mtets_steps_per_s": value * nT / 1e6, "gpu_launches": launches,
            "cg_iterations_per_step_mesh0": iters[0] if iters else [],
        }
        tot_s = sum(m * n for m, n, _ in prof)
        tot_n = sum(n for _, n, _ in prof)
        if tot_n > 0:
            peak, peak_src = measured_peak_gbs()
            mean_s, bytes_launch = tot_s / tot_n, float(np.mean([b for _, _, b in prof]))
            kern = "kb_spmv<1> (q = Keff d + per-mesh d.q over all meshes of a batch context)" if args.group > 0 else "k_spmv_rows3<1>"
            line["roofline"] = {
                "kernel": kern + ", CUDA-event pairs around every 16th PCG iteration's launch inside the timed steps (rank 0)",
                "bound": "hbm", "achieved": bytes_launch / mean_s / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": bytes_launch / mean_s / 1e9 / peak, "traffic": ncu_traffic(f"batch{args.group}x{nx}") if args.group > 0 else None,
                "algorithmic_bytes_per_launch": bytes_launch, "mean_launch_seconds": mean_s, "samples": tot_n,
                "bytes_model": "per context: 8 B/nnz values + 4 B per 3x3 block column + 52 B per block row, all meshes iterating",
                "traffic_source": "profiles/ncu_traffic.json (ncu --set full capture of this kernel; null when none exists for this batch shape)",
            }
        args.emit(line)
    for s_ in sims:
        s_.close()
    if world > 1:
        dist.destroy_process_group()
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
    ap.add_argument("--nx", type=int, default=56, help="cube resolution: 56 = configs[1] (1M tets), 120 = configs[2] (10M tets)")
    ap.add_argument("--cpu-cg-iters", type=int, default=40, help="PCG iterations in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=0, help="configs[3]: step a batch of this many independent meshes (use with --nx 33)")
    ap.add_argument("--streams", type=int, default=8, help="host threads / concurrent contexts per GPU in --batch mode (B200, 32 meshes of 196,608 tets on one GPU: 75.2 / 81.1 / 79.7 mesh-steps/s with 4 / 8 / 16)")
    ap.add_argument("--group", type=int, default=-1, help="--batch mode: meshes per batch context (fb_create_batch: one block-diagonal system, PCG scalars and stopping rule per mesh); -1 = all meshes of the rank in one context (B200, 32 meshes of 196,608 tets: 109 mesh-steps/s), 0 = one context per mesh on --streams host threads (79.6)")
    ap.add_argument("--partitioned", action="store_true",
                    help="N>1: split ONE mesh by row blocks across the ranks (NCCL halo exchange, strong scaling) instead of one mesh per rank")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print("note: timing rules ask for >= 3 warm-up steps", file=sys.stderr)
    args.emit = lambda line: emit(saved_stdout, line)
    if args.impl == "reference":
        return run_reference(args)
    return run_batch(args) if args.batch > 0 else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
