"""Summarise `ncu --page raw --csv` exports (tools/ncu_export.sh) into profiles/:
    python tools/ncu_raw_summary.py <tag> <launches.csv> <full_raw.csv> [<more_raw.csv> ...]
writes profiles/<tag>_launches_summary.csv, profiles/<tag>_ncu_full_summary.json and refreshes profiles/traffic.json
(dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel = what bench.py reports as roofline.traffic)."""
import csv, json, os, sys
from collections import defaultdict
tag, launch_csv, raws = sys.argv[1], sys.argv[2], sys.argv[3:]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
d = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    d[r[ki]][0] += 1; d[r[ki]][1] += v
tot = sum(v[1] for v in d.values())
with open(os.path.join(ROOT, "profiles", "%s_launches_summary.csv" % tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, one batched step (33 frames / 32 pairs); cold-cache, serialised: compare SHARES\n")
    f.write("kernel,launches,total_us,share\n")
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        f.write('"%s",%d,%.1f,%.4f\n' % (k.split("(")[0][-60:], v[0], v[1] / 1e3, v[1] / tot))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "smsp__inst_executed.sum"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res, traffic = [], {}
for raw in raws:
    rr = list(csv.reader(open(raw)))
    h, units = rr[0], rr[1]
    for r in rr[2:]:
        name = r[h.index("Kernel Name")]
        if "at::" in name: continue
        e = {"kernel": name.split("(")[0].split("::")[-1]}
        for k in KEYS:
            if k in h: e[k] = (r[h.index(k)] + " " + units[h.index(k)]).strip()
        res.append(e)
        def num(s): return float(s.split()[0].replace(",", "")) * UNIT.get(s.split()[1], 1)
        short = e["kernel"].split("<")[0]
        traffic.setdefault(short, num(e["dram__bytes_read.sum"]) + num(e["dram__bytes_write.sum"]))
json.dump(res, open(os.path.join(ROOT, "profiles", "%s_ncu_full_summary.json" % tag), "w"), indent=1)
old = {}
tp = os.path.join(ROOT, "profiles", "traffic.json")
if os.path.isfile(tp): old = json.load(open(tp))
old.update(traffic)
json.dump(old, open(tp, "w"), indent=1)
for e in res:
    print("%-28s %12s dram r %-18s w %-18s tensor %-6s tc-smem %-6s lsu-smem %-6s fma %-6s issue %-6s" % (
        e["kernel"][:28], e["gpu__time_duration.sum"], e["dram__bytes_read.sum"], e["dram__bytes_write.sum"],
        e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "-").split()[0][:5],
        e.get("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "-").split()[0][:5],
        e.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "-").split()[0][:5],
        e.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "-").split()[0][:5],
        e.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "-").split()[0][:5]))
