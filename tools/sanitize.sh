#!/bin/bash
# compute-sanitizer over the C-ABI tests: memcheck on the small-shape parity tests (every kernel family:
# encode, K2 forward / gradient incl. the mbarrier rings and stream-K segment reduction, K5 subspace + Jacobi, K6, K7)
# and on the imputation tests (K8), racecheck on the bond-step subset.  Summaries land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout ${T_MEM:-400} $CS --tool memcheck --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not golden and not api" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
timeout ${T_MEM2:-240} $CS --tool memcheck --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_impute.py -m gpu -x -q -k "${K_IMP:-median or mean}" > gpurun_out/r02_sanitizer_memcheck_impute.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck_impute.log
timeout ${T_RACE:-300} $CS --tool racecheck --racecheck-report analysis --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${K_RACE:-teacher_forced or bond_split_cutoff}" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log
for f in memcheck memcheck_impute racecheck; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=|Error:|hazard" gpurun_out/r02_sanitizer_$f.log | sort | uniq -c | sort -rn | head -8; done
