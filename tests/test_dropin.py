"""The Vega-derived adapter classes (include/fembrain_b200_vega_classes.hpp: CudaCorotationalForceModel : ForceModel,
CudaVolumeConservingIntegrator2 : IntegratorBaseSparse) against the REFERENCE's headers, and the drop-in library that
oracle/Makefile builds from the reference's own Deformable.cpp with the INTEGRATION.md substitutions.  CPU part: the
classes compile as C++98 with the oracle's flags next to Vega's headers, every pure virtual is implemented, the drop-in
library loads and reports a missing GPU as an error instead of exiting.  The GPU run is tests/test_dropin_gpu.py."""
import os
import subprocess

import numpy as np
import pytest

from fembrain_b200 import api
from tests import cases

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "fembrain_b200_vega_classes.hpp"
// both classes must be concrete (no pure virtual left) and usable through the reference's base pointers
ForceModel *make_force_model(int nV, const double *x, int nT, const int *t) {
  return new fembrain_b200::CudaCorotationalForceModel(nV, x, nT, t, 1E7, 0.46, 1000);
}
IntegratorBaseSparse *make_integrator(int r, SparseMatrix *M, ForceModel *fm, int nC, int *cd) {
  // the reference's call, argument for argument (src/deformable/Deformable.cpp:208-216)
  return new fembrain_b200::CudaVolumeConservingIntegrator2(r, 0.0333, M, fm, 0, nC, cd, 0.0, 0.01, 1, 1E-6, 8);
}
int step(IntegratorBaseSparse *in, double *f, double *q) {
  in->SetExternalForcesToZero();
  in->SetExternalForces(f);
  int rc = in->DoTimestep();
  in->GetqState(q);
  return rc + (in->GetSystemSolveTime() >= 0.0 ? 0 : 1);
}
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference's headers")
def test_adapter_classes_compile_against_reference_headers_as_cxx98(tmp_path):
    src = tmp_path / "classes.cpp"
    src.write_text(SRC)
    vega = os.path.join(REF, "src", "3rdparty", "vegafem")
    cmd = ["g++", "-std=gnu++98", "-fpermissive", "-w", "-c", "-include", os.path.join(ROOT, "oracle", "ref_prelude.h"),
           "-I", os.path.join(vega, "include"), "-I", os.path.join(REF, "src", "3rdparty"), "-I", os.path.join(ROOT, "include"),
           str(src), "-o", str(tmp_path / "classes.o")]
    subprocess.run(cmd, check=True)
    syms = subprocess.run(["nm", "-C", str(tmp_path / "classes.o")], capture_output=True, text=True, check=True).stdout
    for needed in ("fembrain_b200::CudaCorotationalForceModel::GetForceAndMatrix", "fembrain_b200::CudaCorotationalForceModel::GetTangentStiffnessMatrixTopology",
                   "fembrain_b200::CudaVolumeConservingIntegrator2::DoTimestep", "fembrain_b200::CudaVolumeConservingIntegrator2::setConstrainedDOF"):
        assert needed in syms, needed


def test_dropin_library_exports_the_harness_and_uses_the_c_abi(ref_oracle):
    if not ref_oracle.dropin_available():
        pytest.skip("oracle/_ref/libfembrain_dropin.so not built")
    lib = ref_oracle._LIBS["dropin"]
    dyn = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    for name in ("fbdrop_create", "fbdrop_timestep", "fbdrop_get_state", "fbdrop_pick_vertices"):
        assert name in dyn, name
    und = subprocess.run(["nm", "-D", "--undefined-only", lib], capture_output=True, text=True, check=True).stdout
    for name in ("fb_create_with_materials", "fb_step", "fb_set_state", "fb_get_state", "fb_set_external_forces", "fb_set_fixed_vertices"):
        assert name in und, f"the drop-in build must call {name} of libfembrain_b200.so"
    # the reference's CPU integrator is NOT what this Deformable instantiates
    cxx = subprocess.run(["nm", "-C", os.path.join(os.path.dirname(lib), "obj", "drop_Deformable.o")], capture_output=True, text=True, check=True).stdout
    assert "fembrain_b200::CudaVolumeConservingIntegrator2::DoTimestep" in cxx and "vtable for fembrain_b200::CudaVolumeConservingIntegrator2" in cxx
    assert "U VolumeConservingIntegrator::VolumeConservingIntegrator" not in cxx


def test_dropin_reports_a_missing_gpu_without_exiting(ref_oracle):
    from tests.conftest import has_gpu

    if not ref_oracle.dropin_available():
        pytest.skip("oracle/_ref/libfembrain_dropin.so not built")
    if has_gpu():
        pytest.skip("has a GPU: covered by test_dropin_gpu.py")
    v, t, fixed, _ = cases.cube_case(3)
    with pytest.raises(RuntimeError):
        ref_oracle.RefDeformable(v, t, fixed, lib="dropin")
    assert api.load_library() is not None and np.all(np.isfinite(v))
