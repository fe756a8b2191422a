"""fembrain_b200 — B200-native (sm_100a CUDA) implementation of FemBrain's per-frame soft-tissue
FEM step behind a C ABI (include/fembrain_b200.h).  This package is host-side plumbing only:
`api` binds the shared library with ctypes, `meshes` builds the reference's synthetic inputs,
`build` compiles the library.  There is no CPU or PyTorch compute path."""
from .api import FbParams, FemBrainError, Simulation, comm_unique_id, default_params, load_library, partition_ordering, plan_partition, step_many, tetgen_load, trim_memory, veg_load, veg_save  # noqa: F401
from . import meshes  # noqa: F401

__all__ = ["Simulation", "FbParams", "FemBrainError", "default_params", "load_library", "meshes", "comm_unique_id", "plan_partition", "partition_ordering", "veg_load", "veg_save", "tetgen_load", "trim_memory", "step_many"]
