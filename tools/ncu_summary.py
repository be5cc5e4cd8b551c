"""Turn an `ncu --set full` report into a short text summary (per kernel: duration, launch geometry, registers,
DRAM bytes, L2 sectors, issue utilisation, top stall reasons) for profiles/.  Usage: ncu_summary.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"), ("lts__t_sectors.sum", "L2 sectors"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__inst_executed.sum", "warp instructions"), ("sm__cycles_elapsed.max", "SM cycles"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
]
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    for key, label in want:
        if key in ix:
            print("%-32s %s %s" % (label, r[ix[key]], units[ix[key]]))
    print()
