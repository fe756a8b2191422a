#!/bin/bash
# compute-sanitizer over small cases of every kernel family (GPU box).  Output: gpurun_out/r02_sanitizer.txt
# (copied to profiles/).  memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards (the
# __syncwarp-protected record reuse of fb_assembly.cu, the staging barriers, the mbarrier ring of fb_tma.cu);
# synccheck: barrier misuse.  Small meshes only: the tools slow kernels down ~50-100x.
OUT=gpurun_out/r02_sanitizer.txt
: > $OUT
run() {
  echo "=== compute-sanitizer --tool $1 :: ${*:2}" >> $OUT
  timeout 600 compute-sanitizer --tool "$1" --error-exitcode 9 --print-limit 20 "${@:2}" > gpurun_out/_san.log 2>&1
  rc=$?
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK SUMMARY|Error:|Hazard|passed|failed|smoke ok" gpurun_out/_san.log | tail -12 >> $OUT
  echo "exit code $rc" >> $OUT
}
SMALL='tests/test_parity_gpu.py::test_force_and_stiffness_bit_exact tests/test_parity_gpu.py::test_timestep_parity tests/test_batch_gpu.py tests/test_deformable_gpu.py::test_deformable_timestep_matches_compiled_reference_frames'
run memcheck python -m pytest -x -q -m gpu $SMALL -k "two_tetra or cube7 or slab or cube6 or cube5 or egg or batch"
run racecheck python -c "import __graft_entry__ as g; g.smoke()"
run racecheck python -m pytest -x -q -m gpu tests/test_parity_gpu.py::test_force_and_stiffness_bit_exact -k "cube7 or slab"
run synccheck python -c "import __graft_entry__ as g; g.smoke()"
FEMBRAIN_B200_SPMV=tma run racecheck python -c "import __graft_entry__ as g; g.smoke()"
FEMBRAIN_B200_SPMV=tma run memcheck python -c "import __graft_entry__ as g; g.smoke()"
rm -f gpurun_out/_san.log
cat $OUT
