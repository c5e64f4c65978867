python bench.py --no-impute --no-cpu-baseline > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; python - <<PYEOF
import json
d=json.load(open("gpurun_out/r02_bench8.json"))
st=lambda s:{k:v for k,v in s.items() if k!="note"}
print("value",d["value"],"ms/bond",d["ms_per_bond"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"], st(d["svd_stats"]))
print(d["device_time_breakdown_ms"])
b=d["config_B"]; print("B",b["value"],b["ms_per_bond"],b["e2e"]["value"],b["roofline"]["frac"],b["device_time_breakdown_ms"], st(b["svd_stats"]))
PYEOF
MPST_IMPUTE_DEBUG=1 python tools/impute_prof.py 296 median 2>&1 | tail -3; MPST_IMPUTE_DEBUG=1 python tools/impute_prof.py 296 ITS 2>&1 | tail -2; ncu --set full --clock-control none --import-source on -k regex:impute_kernel -c 1 -o gpurun_out/r02_impute_kernel python tools/impute_prof.py 296 median > gpurun_out/r02_ncu_impute.log 2>&1; tail -1 gpurun_out/r02_ncu_impute.log
