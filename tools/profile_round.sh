#!/bin/bash
# One round's kernel profiles on the GPU box: `ncu --set full` reports go to /tmp (they exceed what gpurun copies back),
# only their text summaries come home: gpurun_out/r2_kernels.md (profiles/tools/summarise.py) and one key-counter file per
# kernel (profiles/tools/ncu_keys.py); the reports named in KEEP are copied as they are.
set -u
export OUT=/tmp/prof
KEEP="${KEEP:-uav_pos policy learn}"
bash tools/profile_all.sh "$@" 2>&1 | tail -3
mkdir -p gpurun_out/r2_keys
python profiles/tools/summarise.py $OUT/prof_*.ncu-rep > gpurun_out/r2_kernels.md 2> gpurun_out/r2_kernels.err
for f in $OUT/prof_*.ncu-rep; do
  b=$(basename $f .ncu-rep)
  ncu -i $f --page raw --csv > /tmp/raw_$b.csv 2>/dev/null && python profiles/tools/ncu_keys.py /tmp/raw_$b.csv > gpurun_out/r2_keys/${b#prof_}_keys.txt
done
for k in $KEEP; do cp $OUT/prof_$k.ncu-rep gpurun_out/ 2>/dev/null; done
ls gpurun_out/r2_keys | wc -l
