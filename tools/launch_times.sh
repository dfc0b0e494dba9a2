#!/bin/bash
# dev tool: per-kernel device times (ncu launch list) of one batched solve; usage: launch_times.sh <tag> [prof_run args]
tag=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/prof_run.py --calls 2 "$@" > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/launches_$tag.csv")) if len(r) > 14 and r[0].isdigit()]
half = len(rows) // 2
for r in rows[half:]:
    print(f"  {r[4].split('(')[0].replace('<unnamed>::','').replace('void ',''):28s} {float(r[14])/1e6:8.3f} ms")
PY
