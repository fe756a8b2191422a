"""Builds fembrain_b200/libfembrain_b200.so (sm_100a only) with nvcc, in-tree.

    python -m fembrain_b200.build [--force]

One object per translation unit.  fb_fem.cu is compiled with -fmad=false: the element / assembly
arithmetic is kept bit-identical to the reference (see fb_element_math.h); the PCG kernels in
fb_pcg.cu use FMA.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libfembrain_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
UNITS = {
    "fb_api.cu": [],
    "fb_setup.cu": [],
    "fb_fem.cu": ["-fmad=false"],
    "fb_assembly.cu": ["-fmad=false"],
    "fb_pcg.cu": [],
    "fb_mg.cu": [],
    "fb_dist.cu": [],
    "fb_batch.cu": [],
    "fb_veg.cu": ["-fmad=false"],
    "fb_deformable.cu": ["-fmad=false"],
}
# measured-and-shelved experiments (csrc/experiments/): built only with --experiments / FEMBRAIN_B200_BUILD_EXPERIMENTS=1;
# the default library carries fb_experiments_off.cu instead (stubs that answer "not available")
EXPERIMENT_UNITS = {os.path.join("experiments", "fb_sym.cu"): [], os.path.join("experiments", "fb_pcg_persistent.cu"): [],
                    os.path.join("experiments", "fb_tma.cu"): [], os.path.join("experiments", "fb_experiments_on.cu"): []}
DEFAULT_ONLY_UNITS = {"fb_experiments_off.cu": []}
HEADERS = ["fb_internal.h", "fb_element_math.h", "fb_pcg_common.cuh", os.path.join("..", "..", "include", "fembrain_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False, experiments: bool | None = None) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if experiments is None:
        experiments = os.environ.get("FEMBRAIN_B200_BUILD_EXPERIMENTS") == "1"
    units = dict(UNITS)
    units.update(EXPERIMENT_UNITS if experiments else DEFAULT_ONLY_UNITS)
    stamp = os.path.join(OBJ, "experiments.on" if experiments else "experiments.off")
    other = os.path.join(OBJ, "experiments.off" if experiments else "experiments.on")
    if os.path.exists(other):   # the flavour changed: relink
        os.remove(other)
        force_link = True
    else:
        force_link = not os.path.exists(stamp)
    open(stamp, "w").close()
    nvcc = _nvcc()
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    logs = []
    for unit, extra in units.items():
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJ, os.path.basename(unit).replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", src, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            logs.append(f"$ {' '.join(cmd)}\n{res.stdout}{res.stderr}")
            if res.returncode != 0:
                sys.stderr.write(logs[-1])
                raise RuntimeError(f"nvcc failed on {unit}")
    if force or force_link or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-lnccl"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        logs.append(f"$ {' '.join(cmd)}\n{res.stdout}{res.stderr}")
        if res.returncode != 0:
            sys.stderr.write(logs[-1])
            raise RuntimeError("link failed")
    if logs:
        with open(os.path.join(OBJ, "build.log"), "w") as fh:
            fh.write("\n".join(logs))
        if verbose:
            print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, experiments=True if "--experiments" in sys.argv else None))
