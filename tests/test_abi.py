"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/fembrain_b200.h declares, and refuses loudly to run without a B200 (no fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

import fembrain_b200 as fb
from fembrain_b200 import api
from tests import cases
from tests.conftest import has_gpu


def declared_functions():
    src = open(api.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", src)))


def test_header_compiles_as_c_and_cpp(tmp_path):
    for comp, std, ext in (("gcc", "-std=c99", "c"), ("g++", "-std=c++11", "cpp")):
        f = tmp_path / f"t.{ext}"
        f.write_text('#include "fembrain_b200.h"\nint main(void){ fb_params p; fb_default_params(&p); return fb_abi_version() == FB_ABI_VERSION ? 0 : 1; }\n')
        subprocess.run([comp, std, "-Wall", "-Werror", "-pedantic", "-I", os.path.dirname(api.HEADER_PATH), "-fsyntax-only", str(f)], check=True)


def test_library_exports_every_declared_symbol():
    lib = fb.load_library()
    names = declared_functions()
    assert len(names) >= 60
    for n in names:
        assert hasattr(lib, n), f"{n} declared in fembrain_b200.h but not exported"
    # and the binding covers them all
    assert set(names) == set(lib._fb_signatures), set(names) ^ set(lib._fb_signatures)


def test_c_program_links_against_the_abi(tmp_path):
    """A plain C host program (what a cgo/JNI/C++ caller does) links and runs against the shared library."""
    src = tmp_path / "host.c"
    src.write_text(
        '#include <stdio.h>\n#include "fembrain_b200.h"\n'
        "int main(void){ fb_params p; fb_default_params(&p); fb_context* c = 0;\n"
        " double x[12] = {-1,0,0, 0,0,-2, 1,0,0, 0,2,-1}; int t[4] = {0,1,2,3}; int fx[1] = {1};\n"
        " int st = fb_create(&c, 4, x, 1, t, 1, fx, &p);\n"
        ' printf("%d %s\\n", st, fb_status_string(st)); if (st == FB_OK) { st = fb_step(c); fb_destroy(c); }\n'
        " return (st == FB_OK || st == FB_ERR_NO_DEVICE) ? 0 : 1; }\n")
    exe = tmp_path / "host"
    libdir = os.path.dirname(api.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-I", os.path.dirname(api.HEADER_PATH), str(src), "-o", str(exe), "-L", libdir,
                    "-lfembrain_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.split()[0] in ("0", "2")


def test_defaults_are_the_reference_constants():
    p = fb.default_params()
    assert (p.youngs_modulus, p.poisson_ratio, p.density) == (1e7, 0.46, 1000.0)  # Deformable.cpp:178
    assert p.timestep == 0.0333 and p.damping_mass == 0.0 and p.damping_stiffness == 0.01  # :107-113
    assert p.cg_epsilon == 1e-6 and p.cg_max_iterations == 10000  # PS_VolumeConservingIntegrator.cpp:196-197
    assert p.polar_tolerance == 1e-6 and p.internal_force_scaling == 1.0
    assert fb.load_library().fb_abi_version() == 1


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    v, t, fixed, _ = cases.cube_case(3)
    with pytest.raises(fb.FemBrainError) as e:
        fb.Simulation(v, t, fixed)
    assert e.value.status == api.FB_ERR_NO_DEVICE


def test_invalid_arguments_are_status_codes_not_crashes():
    lib = fb.load_library()
    h = ctypes.c_void_p()
    assert lib.fb_create(ctypes.byref(h), -1, None, 0, None, 0, None, None) == api.FB_ERR_INVALID_ARGUMENT
    assert lib.fb_create(None, 0, None, 0, None, 0, None, None) == api.FB_ERR_INVALID_ARGUMENT
    assert lib.fb_step(None) == api.FB_ERR_INVALID_ARGUMENT
    assert lib.fb_num_dofs(None) == 0
    lib.fb_destroy(None)
    assert b"argument" in lib.fb_status_string(1)


def test_batch_entry_points_validate_before_touching_the_device():
    """fb_create_batch checks counts, mesh-local vertex ids and fixed-vertex lists on the host, so these fail the same way
    with or without a GPU (and never reach fb_create_local)."""
    import numpy as np

    v, t, fixed, _ = cases.cube_case(3)
    with pytest.raises(fb.FemBrainError) as e:
        fb.Simulation(batch=[(v, t, fixed), (v, t + 1000, fixed)])
    assert e.value.status == api.FB_ERR_BAD_MESH and "mesh 1" in str(e.value)
    with pytest.raises(fb.FemBrainError) as e:
        fb.Simulation(batch=[(v, t, np.array([0, 0], np.int32))])
    assert e.value.status == api.FB_ERR_INVALID_ARGUMENT and "twice" in str(e.value)
    with pytest.raises(fb.FemBrainError) as e:
        fb.Simulation(batch=[(v, t, np.array([len(v)], np.int32))])
    assert e.value.status == api.FB_ERR_INVALID_ARGUMENT
    lib = fb.load_library()
    h = ctypes.c_void_p()
    assert lib.fb_create_batch(ctypes.byref(h), 0, None, None, None, None, None, None, None) == api.FB_ERR_INVALID_ARGUMENT
    assert lib.fb_batch_count(None) == 0
    assert lib.fb_batch_offsets(None, None, None) == api.FB_ERR_INVALID_ARGUMENT
    assert lib.fb_batch_last_cg_iterations(None, None, None) == api.FB_ERR_INVALID_ARGUMENT


def test_loading_the_library_before_torch_does_not_break_torch():
    """Both link libnccl.so.2; the first copy mapped wins.  api.load_library maps the PyTorch wheel's copy first, so a process
    that touches fembrain_b200 before `import torch` (tests/test_fullsize_gpu.py run alone did) still imports torch."""
    import subprocess
    import sys

    code = "import fembrain_b200.api as a; a.load_library(); import torch; print('ok', torch.__version__)"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stderr[-800:]
