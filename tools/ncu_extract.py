"""Key counters of every kernel in one or more `ncu --set full` reports as one CSV (the table kept under profiles/).
    python tools/ncu_extract.py gpurun_out/x/*.ncu-rep > profiles/<name>.csv"""
import csv
import io
import subprocess
import sys

COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
w = csv.writer(sys.stdout)
first = True
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) if c in hdr else None for c in COLS]
    if first:
        w.writerow(["report"] + COLS)
        w.writerow([""] + [units[i] if i is not None else "" for i in idx])
        first = False
    for r in rows[2:]:
        w.writerow([rep.split("/")[-1]] + [(r[i][:90] if i is not None else "") for i in idx])
