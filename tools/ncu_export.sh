#!/bin/bash
# Runs on the GPU box: ncu captures of one batched step, exported to CSV there (the .ncu-rep files with sources are too big to
# travel back whole).  usage: tools/ncu_export.sh <tag> [source-kernel-regex] [--lean]
#   --lean: launch lists + one --set full pass of the ring-image step + the source page of the regex'd kernels only
set -u
TAG=${1:-r2}; SRC=${2:-"respond_score|conv3_tc|conv12_"}; LEAN=${3:-}
O=gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv python tools/ncu_step.py > /dev/null 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_scans.csv python tools/ncu_step.py --scans > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none -o /tmp/${TAG}_full python tools/ncu_step.py > $O/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_full_raw.csv 2>/dev/null
if [ "$LEAN" != "--lean" ]; then
ncu --profile-from-start off --set full --clock-control none -k regex:"scan_brick|ring_" -o /tmp/${TAG}_full_scans python tools/ncu_step.py --scans >> $O/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}_full_scans.ncu-rep --page raw --csv > $O/${TAG}_full_scans_raw.csv 2>/dev/null
fi
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$SRC" -o /tmp/${TAG}_src python tools/ncu_step.py >> $O/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}_src.ncu-rep --page source --csv > $O/${TAG}_source.csv 2>/dev/null
ls -la /tmp/${TAG}_*.ncu-rep $O | tail -12
gzip -f $O/${TAG}_source.csv
