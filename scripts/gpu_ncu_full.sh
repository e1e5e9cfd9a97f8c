# ncu --set full captures of the dominant kernel of each workload (one GPU). usage: bash scripts/gpu_ncu_full.sh [tag]
TAG=${1:-r01}
cap() {  # workload, kernel regex, skip, extra bench args
  local w=$1 k=$2 s=$3; shift 3
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu "$@" > gpurun_out/plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 2 -f -o gpurun_out/prof_${TAG}_$w \
      python bench.py --workload $w --steps 6 --warmup 3 --no-cpu "$@" > gpurun_out/ncu_$w.log 2>&1
  echo "ncu $w exit $?"; tail -1 gpurun_out/ncu_$w.log
  # summarise on the box (gpurun copies back at most 64 MiB): keep the text, drop the report unless KEEP_REP names it
  python scripts/ncu_summary.py gpurun_out/prof_${TAG}_$w.ncu-rep > gpurun_out/${TAG}_ncu_full_$w.txt 2>/dev/null
  case " $KEEP_REP " in *" $w "*) ;; *) rm -f gpurun_out/prof_${TAG}_$w.ncu-rep ;; esac
}
WL=${WORKLOADS:-"c2 c1 c3_hopper c3_halfcheetah c4 c4_rollout rollout rollout_rec"}
for w in $WL; do
  case $w in
    c2) cap c2 cartpole_step_f32_tma 2 ;;
    c1) cap c1 cartpole_step_f32_small 2 ;;
    c3_hopper) cap c3_hopper reward_terminal 1 ;;
    c3_halfcheetah) cap c3_halfcheetah reward_terminal 1 ;;
    c4) cap c4 charged_ball_step 1 ;;
    c4_rollout) cap c4_rollout rollout_f32_kernel 1 --total-log2 24 ;;
    rollout) cap rollout rollout_f32_kernel 1 ;;
    rollout_rec) cap rollout_rec rollout_f32_kernel 1 ;;
  esac
done
