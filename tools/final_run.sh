#!/bin/bash
# Round-end validation on the GPU box: smoke, GPU tests, both bench arms, launch list and one ncu --set full capture.
# Everything lands in gpurun_out/ (copied into profiles/ afterwards).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 300 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_final.csv \
    python bench.py --steps 1 --warmup 0 --series-length 6 --no-cpu-baseline --no-impute > /dev/null 2>&1
timeout 280 ncu --set full --clock-control none --import-source on -k regex:bond_grad_kr_kernel -s 30 -c 1 \
    -o gpurun_out/r01_bond_grad_kr_final python bench.py --no-cpu-baseline --no-impute --steps 1 --warmup 0 --series-length 24 > gpurun_out/ncu_grad.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print("clocks", d["clocks"], "cpu", d["cpu_baseline"]["value"], "impute", d["impute"]["value"], "frac", d["roofline"]["frac"])
r = json.load(open("gpurun_out/bench_reference.json"))
print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
ls -la gpurun_out | grep -E "final|bench_"
