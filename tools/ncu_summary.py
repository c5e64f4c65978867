"""Summarise an `ncu --set full` report (one kernel launch per report id) into the text kept under profiles/.
    python tools/ncu_summary.py report.ncu-rep [title] > profiles/<name>.txt
Reads the report with `ncu -i ... --page raw --csv`; keeps duration, pipe utilisation, DRAM / L2 traffic, occupancy,
issue statistics and the top warp-stall reasons."""
import csv, subprocess, sys
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
KEEP = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__sass_inst_executed_op_shared_ld.sum",
        "sm__sass_inst_executed_op_shared_st.sum"]
print(f"# {title}")
print(f"# source: {rep}   (ncu --set full --clock-control none; one launch per block below)")
for r in rows[2:]:
    d = dict(zip(h, r))
    print(f"\n== {d.get('Kernel Name', '?')[:110]}")
    for k in KEEP:
        if k in d and d[k] != "":
            print(f"  {k:80s} {d[k]:>16s} {u[h.index(k)]}")
    stalls = [(float(d[k]), k) for k in h if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and d.get(k)]
    for v, k in sorted(stalls, reverse=True)[:6]:
        print(f"  stall {k:74s} {v:16.3f}")
