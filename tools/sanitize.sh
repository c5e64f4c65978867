#!/bin/bash
# compute-sanitizer over the C-ABI tests: memcheck on the small-shape parity / encoding / imputation tests,
# racecheck + synccheck on the bond-step and imputation subsets (the mbarrier rings, the named-barrier free Jacobi, the
# stream-K segment reduction).  Summaries land in gpurun_out/ (copied to profiles/ afterwards).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SEL_MEM="tests/test_gpu_parity.py tests/test_gpu_encodings.py tests/test_gpu_impute.py"
SEL_RACE="tests/test_gpu_parity.py -k bond_step_or_loss_grad_or_overlaps_or_split"
timeout ${T_MEM:-900} $CS --tool memcheck --error-exitcode 0 --print-limit 20 python -m pytest $SEL_MEM -m gpu -x -q > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
timeout ${T_RACE:-900} $CS --tool racecheck --racecheck-report analysis --error-exitcode 0 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${K_RACE:-bond or grad or split or sweep}" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log
timeout ${T_SYNC:-600} $CS --tool synccheck --error-exitcode 0 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${K_RACE:-bond or grad or split or sweep}" > gpurun_out/r02_sanitizer_synccheck.log 2>&1
echo "synccheck rc=$?" >> gpurun_out/r02_sanitizer_synccheck.log
for f in memcheck racecheck synccheck; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/r02_sanitizer_$f.log | tail -5; done
