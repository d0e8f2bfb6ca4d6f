#!/bin/bash
# One `ncu --set full` capture of the dominant kernel of every workload (run on the GPU box, after the same commands
# exited 0 without ncu).  Reports land in gpurun_out/prof_<name>.ncu-rep; profiles/tools/summarise.py turns them into
# profiles/<round>/kernels.md.
set -u
out=${OUT:-gpurun_out}; mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on -f"
prof_env() { # workload kernel-regex
  python bench.py --workload $1 --steps 12 --warmup 3 --no-extras > $out/plain_$1.log 2>&1 || { echo "plain $1 failed"; return; }
  $NCU -k regex:$2 -s 4 -c 1 -o $out/prof_$1 python bench.py --workload $1 --steps 12 --warmup 3 --no-extras > $out/ncu_$1.log 2>&1
}
prof_micro() { # kind kernel-regex
  python tools/microbench.py $1 4 > $out/plain_$1.log 2>&1 || { echo "plain $1 failed"; return; }
  $NCU -k regex:$2 -s 3 -c 1 -o $out/prof_$1 python tools/microbench.py $1 4 > $out/ncu_$1.log 2>&1
}
for w in "$@"; do
  case $w in
    uav_pos) prof_env uav_pos uav_pos_step ;;
    uav_att) prof_env uav_att uav_att_step ;;
    cartpole) prof_env cartpole cartpole_step ;;
    ugvo) prof_env ugvo ugvo_step ;;
    uavr_hover) prof_env uavr_hover uavrobust_step ;;
    soi|fas|fas_discrete|ballbalancer|twolink|ugv) prof_env $w env_step_kernel ;;
    gae) prof_micro gae gae_kernel ;;
    gae_flags) prof_micro gae_flags gae_kernel ;;
    mc_returns) prof_micro mc_returns mc_returns_kernel ;;
    policy) prof_micro policy policy_umma16 ;;
    policy_wide) prof_micro policy_wide policy_umma_kernel ;;
    learn) MB=262144 prof_micro learn ppo2_grad ;;
    learn_small) MB=4096 python tools/microbench.py learn 4 > $out/plain_learn_small.log 2>&1 && MB=4096 $NCU -k regex:ppo2_grad -s 3 -c 1 -o $out/prof_learn_small python tools/microbench.py learn 4 > $out/ncu_learn_small.log 2>&1 ;;
    adam) MB=4096 $NCU -k regex:adam_kernel -s 3 -c 1 -o $out/prof_adam python tools/microbench.py learn 4 > $out/ncu_adam.log 2>&1 ;;
    gae_scan) GAE_N=8 prof_micro gae_small_scan gae_scan_kernel ;;
    ugvo_reset) $NCU -k regex:ugvo_autoreset -s 4 -c 1 -o $out/prof_ugvo_reset python bench.py --workload ugvo --steps 12 --warmup 3 --no-extras > $out/ncu_ugvo_reset.log 2>&1 ;;
    cartpole_rollout) python tools/rollout_once.py cartpole 65536 200 > $out/plain_cartpole_rollout.log 2>&1 && $NCU -k regex:cartpole_rollout -c 1 -o $out/prof_cartpole_rollout python tools/rollout_once.py cartpole 65536 200 > $out/ncu_cartpole_rollout.log 2>&1 ;;
    policy_fp32) prof_micro policy_fp32 policy_forward_kernel ;;
    norm_stats) python tools/microbench.py norm 4 > $out/plain_norm.log 2>&1 && $NCU -k regex:norm_batch_stats -s 3 -c 1 -o $out/prof_norm_stats python tools/microbench.py norm 4 > $out/ncu_norm_stats.log 2>&1 ;;
    norm_apply) $NCU -k regex:norm_merge_apply -s 3 -c 1 -o $out/prof_norm_apply python tools/microbench.py norm 4 > $out/ncu_norm_apply.log 2>&1 ;;
  esac
done
ls -la $out/prof_*.ncu-rep | wc -l
