#!/usr/bin/env bash
# Text extract of an ncu report (runs without a GPU):  bash tools/ncu_extract.sh gpurun_out/r02_stream_kernel.ncu-rep > profiles/r02_cfg3_stream_kernel_ncu.txt
# Also prints the per-launch DRAM traffic that bench.py reads from profiles/traffic.json ("cfg3_n1").
set -eu
REP="$1"
CSV="${REP%.ncu-rep}.raw.csv"
ncu -i "$REP" --page raw --csv > "$CSV"
python - "$CSV" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled", "sm__cycles_active", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct")
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k) for k in keep) or h in ("Kernel Name",):
        print(f"{h:90s} {u:12s} {v}")
def num(name):
    return float(vals[hdr.index(name)].replace(",", "")) if name in hdr else float("nan")
def to_bytes(name):
    u = units[hdr.index(name)].lower() if name in hdr else "byte"
    return num(name) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
print("\ntraffic (dram read + write, bytes per launch):", to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
PY
