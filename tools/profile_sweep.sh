#!/bin/bash
# Full ncu sweep with the summaries made ON the GPU box: the .ncu-rep files (5 MB each) exceed what gpurun copies back,
# so only text goes into gpurun_out/ (kernels_sweep.md, per-kernel op histograms / per-line tables for the headliners).
set -u
out=gpurun_out
mkdir -p $out /tmp/reps
all="uav_pos uav_att uavr_hover cartpole ugvo soi fas fas_discrete ballbalancer twolink ugv gae gae_flags mc_returns norm_stats norm_apply policy policy_fp32"
for w in $all; do
  tools/profile_all.sh $w > /dev/null 2>&1
  mv $out/prof_*.ncu-rep /tmp/reps/ 2>/dev/null
done
python profiles/tools/summarise.py $(for w in $all; do echo /tmp/reps/prof_$w.ncu-rep; done) > $out/kernels_sweep.md 2> $out/kernels_sweep.err
for w in uav_pos uav_att cartpole policy; do
  ncu -i /tmp/reps/prof_$w.ncu-rep --page source --csv > /tmp/reps/$w.src.csv 2>/dev/null
  python profiles/tools/op_hist.py /tmp/reps/$w.src.csv $([ $w = cartpole ] && echo 65536 || echo 1048576) > $out/${w}_ophist.txt 2>&1
  ncu -i /tmp/reps/prof_$w.ncu-rep --page source --csv --print-source sass,cuda > /tmp/reps/$w.cuda.csv 2>/dev/null
  python profiles/tools/lines.py /tmp/reps/$w.cuda.csv 40 > $out/${w}_lines.txt 2>&1
  ncu -i /tmp/reps/prof_$w.ncu-rep --page raw --csv > /tmp/reps/$w.raw.csv 2>/dev/null
  python profiles/tools/ncu_keys.py /tmp/reps/$w.raw.csv > $out/${w}_keys.txt 2>&1
done
rm -f $out/plain_*.log $out/ncu_*.log
ls -la $out | head -30
