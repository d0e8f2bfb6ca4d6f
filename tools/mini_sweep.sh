#!/bin/bash
# per-source-line tables for a few workloads, summarised on the box: tools/mini_sweep.sh <workload...>
set -u
out=gpurun_out; mkdir -p /tmp/reps
for w in "$@"; do
  tools/profile_all.sh $w > /dev/null 2>&1
  mv $out/prof_$w.ncu-rep /tmp/reps/
  ncu -i /tmp/reps/prof_$w.ncu-rep --page source --csv --print-source sass,cuda > /tmp/reps/$w.cuda.csv 2>/dev/null
  python profiles/tools/lines.py /tmp/reps/$w.cuda.csv 30 > $out/${w}_lines.txt 2>&1
done
rm -f $out/plain_*.log $out/ncu_*.log
