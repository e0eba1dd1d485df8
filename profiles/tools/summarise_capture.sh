#!/bin/bash
# Turn one `ncu --set full --import-source on` capture into the summaries committed under profiles/:
#   summarise_capture.sh gpurun_out/r2_step_c2.ncu-rep profiles/r2_step_c2
# -> <prefix>_details.txt, _raw_metrics.csv (selected metrics + stall sample counts), _functions.txt,
#    _source_hotspots.txt, _stalls_no_instruction.txt
set -e
REP=$1; PRE=$2; HERE=$(dirname "$0")
ncu -i "$REP" --page details > "${PRE}_details.txt" 2>/dev/null
ncu -i "$REP" --page raw --csv > /tmp/_raw.csv 2>/dev/null
ncu -i "$REP" --page source --csv --print-source cuda,sass > /tmp/_src.csv 2>/dev/null
python - "$PRE" <<'PY'
import csv, sys
pre = sys.argv[1]
rows = list(csv.reader(open('/tmp/_raw.csv')))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed.ratio"]
idx = [hdr.index(w) for w in want if w in hdr]
with open(pre + "_raw_metrics.csv", "w") as f:
    f.write(",".join(hdr[i] for i in idx) + "\n")
    f.write(",".join(units[i] for i in idx) + "\n")
    f.write(",".join(vals[i].replace(",", ";") for i in idx) + "\n")
    for i, h in enumerate(hdr):
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            f.write(f"# {h} {vals[i]}\n")
PY
python "$HERE/ncu_functions.py" /tmp/_src.csv > "${PRE}_functions.txt"
python "$HERE/ncu_lines.py" /tmp/_src.csv 2>/dev/null | head -42 | cut -c1-190 > "${PRE}_source_hotspots.txt"
python "$HERE/ncu_stalls.py" /tmp/_src.csv stall_no_inst 25 > "${PRE}_stalls_no_instruction.txt"
ls -la ${PRE}_*
