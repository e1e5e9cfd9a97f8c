# ncu --set full captures of the dominant kernel of each workload (one GPU). usage: bash scripts/gpu_ncu_full.sh [tag]
TAG=${1:-r01}
cap() {  # workload, kernel regex, skip
  local w=$1 k=$2 s=$3
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu > gpurun_out/plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 2 -f -o gpurun_out/prof_${TAG}_$w \
      python bench.py --workload $w --steps 6 --warmup 3 --no-cpu > gpurun_out/ncu_$w.log 2>&1
  echo "ncu $w exit $?"; tail -2 gpurun_out/ncu_$w.log
}
cap c2 cartpole_step_f32 1
cap c3_hopper reward_terminal 1
cap c3_halfcheetah reward_terminal 1
cap c4 charged_ball_step 1
