"""The header-only C++ adapter (include/fembrain_b200_vega.hpp) must compile as C++98 — the dialect the
reference itself needs — and link against the C-ABI library; on a GPU box the program
also runs one DoTimestep through the reference-shaped interface."""
import os
import subprocess

from fembrain_b200 import api

SRC = r'''
#include <stdio.h>
#include <vector>
#define FEMBRAIN_B200_NO_EXIT
#include "fembrain_b200_vega.hpp"
int main() {
  // VolMeshSamples::CreateTwoTetra (the mesh main.cpp:833 builds), vertex 0 fixed
  double verts[15] = {-1,0,0, 1,0,0, 0,0,-1, 0,0,1, 0,2,0};
  int elements[8] = {0,2,3,4, 1,2,3,4};
  int fixedDofs[3] = {0,1,2};
  int dof = 15;
  try {
    fembrain_b200::CudaVolumeConservingIntegrator integrator(5, verts, 2, elements, 1E7, 0.46, 1000, dof, 0.0333, 0, 3, fixedDofs,
                                                             0.0, 0.01, 1, 1E-6, 8);
    std::vector<double> ext(dof, 0.0), q(dof), qv(dof), qa(dof);
    ext[3 * 4 + 0] = 1e4;
    integrator.SetExternalForcesToZero();
    integrator.SetExternalForces(&ext[0]);
    if (integrator.DoTimestep() != 0) return 3;
    integrator.GetqState(&q[0], &qv[0], &qa[0]);
    integrator.SetqState(&q[0], &qv[0], &qa[0]);
    printf("RAN r=%d it=%d q=%.17g solve=%g\n", integrator.Getr(), integrator.GetLastCGIterations(), q[12], integrator.GetSystemSolveTime());
    return (q[0] == 0.0 && q[12] > 0.0) ? 0 : 4;
  } catch (int) {
    printf("NODEVICE %s\n", fb_last_error_string());
    return 0;
  }
}
'''


def test_adapter_compiles_as_cxx98_and_links(tmp_path):
    src = tmp_path / "host.cpp"
    src.write_text(SRC)
    exe = tmp_path / "host"
    libdir = os.path.dirname(api.LIB_PATH)
    subprocess.run(["g++", "-std=gnu++98", "-Wall", "-I", os.path.dirname(api.HEADER_PATH), str(src), "-o", str(exe), "-L", libdir,
                    "-lfembrain_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    last = out.strip().splitlines()[-1]
    assert last.startswith("RAN") or last.startswith("NODEVICE"), out


import pytest  # noqa: E402


@pytest.mark.gpu
def test_adapter_program_runs_on_the_gpu_and_matches_the_ctypes_path(tmp_path):
    """The same C++98 host program, executed on the B200: DoTimestep through the reference-shaped class must give the bits the
    ctypes path gives (same C ABI underneath)."""
    import numpy as np

    import fembrain_b200 as fb

    src = tmp_path / "host.cpp"
    src.write_text(SRC)
    exe = tmp_path / "host"
    libdir = os.path.dirname(api.LIB_PATH)
    subprocess.run(["g++", "-std=gnu++98", "-Wall", "-I", os.path.dirname(api.HEADER_PATH), str(src), "-o", str(exe), "-L", libdir,
                    "-lfembrain_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    last = out.strip().splitlines()[-1]
    assert last.startswith("RAN"), out
    fields = dict(kv.split("=") for kv in last.split()[1:])
    verts = np.array([[-1, 0, 0], [1, 0, 0], [0, 0, -1], [0, 0, 1], [0, 2, 0]], dtype=np.float64)
    tets = np.array([[0, 2, 3, 4], [1, 2, 3, 4]], dtype=np.int32)
    sim = fb.Simulation(verts, tets, constrained_dofs=[0, 1, 2])
    f = np.zeros(15)
    f[12] = 1e4
    sim.set_external_forces(f)
    sim.do_timestep()
    q = sim.get_state()[0]
    assert float(fields["q"]) == q[12] and int(fields["it"]) == sim.last_cg_iterations and int(fields["r"]) == 15
