"""Key metrics of the first kernel of an .ncu-rep as text (and, with --json, the DRAM traffic record bench.py reads).
Usage: python profiles/ncu_summary.py <rep> [--json out.json --kernel NAME --bytes-per-gpu N --source TEXT]"""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, r = rows[0], rows[1], rows[2]
m = {n: (r[i], units[i]) for i, n in enumerate(h)}
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
for k in want:
    if k in m:
        print("%-84s %s %s" % (k, m[k][0], m[k][1]))


def num(k):
    v, u = m[k]
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


if "--json" in sys.argv:
    a = sys.argv
    rec = {"kernel": a[a.index("--kernel") + 1], "bytes_per_gpu": int(a[a.index("--bytes-per-gpu") + 1]),
           "dram_bytes_per_launch": int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum")),
           "dram_bytes_read": int(num("dram__bytes_read.sum")), "dram_bytes_write": int(num("dram__bytes_write.sum")),
           "source": a[a.index("--source") + 1]}
    json.dump(rec, open(a[a.index("--json") + 1], "w"), indent=1)
    print(json.dumps(rec))
