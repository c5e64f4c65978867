#!/bin/bash
# Two-rank validation on a 2-GPU box: the multi-rank parity test and the strong-scaling bench line (gpurun --gpus 2).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/r02_multi_2rank_pytest.log 2>&1; tail -3 gpurun_out/r02_multi_2rank_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --impute-instances 4096 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -2 gpurun_out/r02_bench_n2.err
python - <<PYEOF
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_n2.json").read().splitlines() if l.startswith("{")][-1])
st=lambda s:{k:v for k,v in s.items() if k!="note"}
print("N=2 value",d["value"],"ms/bond",d["ms_per_bond"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"], st(d["svd_stats"]))
print(d["device_time_breakdown_ms"])
print("impute", d.get("impute",{}).get("value"))
PYEOF
