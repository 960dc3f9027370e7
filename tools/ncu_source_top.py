"""Top SASS instructions by stall samples from `ncu --page source --csv` exports (tools/ncu_export.sh):
    python tools/ncu_source_top.py gpurun_out/<tag>_source.csv.gz <kernel substring> [N]"""
import csv, gzip, sys
from collections import Counter
path, want = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
op = gzip.open if path.endswith(".gz") else open
rows = list(csv.reader(op(path, "rt")))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name" and want in rows[i][1]:
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr): body.append(rows[j])
            j += 1
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[si]) for r in body)
        print(rows[i][1][:100], "instructions", len(body), "samples", tot, "warp-instr executed", sum(int(r[ii]) for r in body))
        cat = Counter()
        for r in body:
            for k in stall: cat[hdr[k]] += int(r[k])
        print("  by stall:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(1, sum(cat.values()))) for k, v in cat.most_common(9)))
        order = sorted(range(len(body)), key=lambda k: -int(body[k][si]))[:topn]
        for k in sorted(order):
            r = body[k]
            top = max(stall, key=lambda c: int(r[c]))
            print("  %5d %6.2f%% x%-8s %-60s %s" % (k, 100.0 * int(r[si]) / max(1, tot), r[ii], r[1].strip()[:60], hdr[top][6:]))
        i = j
    else:
        i += 1
