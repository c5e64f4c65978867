timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-impute > gpurun_out/r02_bench10.json 2> gpurun_out/r02_bench10.err; tail -3 gpurun_out/r02_bench10.err; python - <<PYEOF
import json
txt=open("gpurun_out/r02_bench10.json").read()
d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
st=lambda s:{k:v for k,v in s.items() if k!="note"}
print("value",d["value"],"ms/bond",d["ms_per_bond"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"], st(d["svd_stats"]))
print(d["device_time_breakdown_ms"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["seconds"])
b=d["config_B"]; print("B",b["value"],b["ms_per_bond"],b["e2e"]["value"],b["roofline"]["frac"],b["device_time_breakdown_ms"], st(b["svd_stats"]))
PYEOF
