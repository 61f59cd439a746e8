"""Selected metrics of every launch in an .ncu-rep as CSV (run where ncu is installed):
    python profiles/tools/ncu_summary.py gpurun_out/r2_full.ncu-rep [name-regex] > profiles/r2_ncu_....csv"""
import csv
import io
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
           "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units, body = rows[0], rows[1], rows[2:]
cols = [head.index("Kernel Name")] + [i for i, h in enumerate(head) if h in METRICS]
w = csv.writer(sys.stdout)
w.writerow([head[i] for i in cols])
w.writerow([units[i] for i in cols])
for r in body:
    if pat is None or pat.search(r[cols[0]]):
        w.writerow([r[i] for i in cols])
