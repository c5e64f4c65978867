#!/bin/bash
# Round-end validation on the GPU box: smoke, GPU tests, both bench arms, launch list, ncu --set full captures of the
# dominant kernels, racecheck.  Everything lands in gpurun_out/ (summaries are copied into profiles/ afterwards).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r02
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err
tail -c 300 gpurun_out/${R}_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
SHORT="bench.py --steps 1 --warmup 0 --samples 262144 --series-length 24 --no-cpu-baseline --no-impute --no-config-b"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_final.csv python $SHORT > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bond_grad_kr_kernel -s 40 -c 1 -o gpurun_out/${R}_grad_kr_final python $SHORT > gpurun_out/ncu_grad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:krao_slab_kernel -s 60 -c 1 -o gpurun_out/${R}_krao_slab_final python $SHORT > gpurun_out/ncu_krao.log 2>&1
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "teacher_forced or bond_split_cutoff" > gpurun_out/${R}_sanitizer_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${R}_sanitizer_racecheck.log | tail -3
python - <<'PY'
import json
def last_json(p):
    return json.loads([l for l in open(p).read().splitlines() if l.startswith("{")][-1])
d = last_json("gpurun_out/r02_bench_default.json")
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print("clocks", d["clocks"], "cpu", d["cpu_baseline"]["value"], "impute", d["impute"]["value"], "frac", d["roofline"]["frac"])
print("B", d["config_B"]["value"], d["config_B"]["roofline"]["frac"], d["config_B"].get("api_fitMPS", {}).get("value"))
r = last_json("gpurun_out/r02_bench_reference.json")
print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
ls -la gpurun_out | grep -E "final|bench_"
