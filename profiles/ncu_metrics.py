"""Print the headline metrics of every launch in an .ncu-rep (read with `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max",
    "sm__pipe_tensor_subpipe_utc_cycles_active",
    "sm__pipe_tensor_cycles_active",
    "sm__inst_executed_pipe_tensor",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "smsp__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum",
]

rep = sys.argv[1]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("====", name[:90])
    for i, h in enumerate(hdr):
        if any(k in h for k in KEYS):
            print(f"  {h} = {r[i]} {units[i]}")
