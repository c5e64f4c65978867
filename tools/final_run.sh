#!/bin/bash
# Round-end validation on the GPU box: smoke, GPU tests, the default bench line, the split-only micro-sweep and a
# bounded memcheck of the code added last (two-pass split, K8 table mode).  Everything lands in gpurun_out/ (summaries are
# copied into profiles/ afterwards).  The ncu captures / launch lists / racecheck of this round were taken by the
# earlier version of this script (profiles/r02_*_ncu.txt, r02_launch_list_final.txt, r02_sanitizer_*.txt).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r02
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${R}_gpu_pytest_final.log
timeout 600 python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err
tail -c 300 gpurun_out/${R}_bench_default.err
timeout 120 python tools/bench_configs.py Esplits > /dev/null 2> gpurun_out/${R}_configE_splits.err
tail -2 gpurun_out/${R}_configE_splits.err | cut -c1-400
SAN="/usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 0 --print-limit 5"
timeout 70 $SAN python -m pytest tests/test_gpu_zz_wide_links.py -m gpu -x -q -k "rank_deficient and 70" > gpurun_out/${R}_sanitizer_memcheck_twopass.log 2>&1
timeout 70 $SAN python -m pytest tests/test_gpu_zz_impute_tables.py -m gpu -x -q -k "SLTD and backwards" > gpurun_out/${R}_sanitizer_memcheck_impute_tables.log 2>&1
grep -hE "ERROR SUMMARY|passed|failed" gpurun_out/${R}_sanitizer_memcheck_twopass.log gpurun_out/${R}_sanitizer_memcheck_impute_tables.log | tail -4
python - <<'PY'
import json
def last_json(p):
    return json.loads([l for l in open(p).read().splitlines() if l.startswith("{")][-1])
d = last_json("gpurun_out/r02_bench_default.json")
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print("clocks", d["clocks"], "cpu", d["cpu_baseline"]["value"], "impute", d["impute"]["value"], "frac", d["roofline"]["frac"])
print("B", d["config_B"]["value"], d["config_B"]["roofline"]["frac"], d["config_B"].get("api_fitMPS", {}).get("value"))
PY
