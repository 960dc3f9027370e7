"""Summarise ncu artefacts into profiles/: python tools/ncu_summary.py <launches.csv> <full.ncu-rep> <tag>"""
import csv, json, subprocess, sys
from collections import defaultdict
launch_csv, rep, tag = sys.argv[1:4]
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
d = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    d[r[ki]][0] += 1; d[r[ki]][1] += v
tot = sum(v[1] for v in d.values())
with open("profiles/%s_launches_summary.csv" % tag, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    f.write("kernel,launches,total_us,share\n")
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        f.write('"%s",%d,%.1f,%.4f\n' % (k, v[0], v[1] / 1e3, v[1] / tot))
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines())); hdr = rows[0]; units = rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
res = []
for r in rows[2:]:
    res.append({k: (r[hdr.index(k)] + " " + units[hdr.index(k)]).strip() for k in keys if k in hdr})
json.dump(res, open("profiles/%s_ncu_full_summary.json" % tag, "w"), indent=1)
traffic = {}
for r in res:
    name = r["Kernel Name"].split("(")[0].split("::")[-1].split("<")[0].strip()
    def num(s): return float(s.split()[0].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[s.split()[1]]
    traffic.setdefault(name, num(r["dram__bytes_read.sum"]) + num(r["dram__bytes_write.sum"]))
    print("%-40s %10s  dram r %-16s w %-16s tensor %-10s fma %-10s issue %-8s warps %s" % (
        name[:40], r["gpu__time_duration.sum"], r["dram__bytes_read.sum"], r["dram__bytes_write.sum"],
        r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "-").split()[0],
        r.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "-").split()[0],
        r.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "-").split()[0],
        r.get("sm__warps_active.avg.pct_of_peak_sustained_active", "-").split()[0]))
json.dump(traffic, open("profiles/traffic.json", "w"), indent=1)
