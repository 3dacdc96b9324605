#!/bin/bash
# One GPU visit: every step under its own timeout so that a hang costs minutes, not the visit.
# usage: tools/gpu_round.sh TAG [steps...]   steps: tests tf32 smoke bench ref search2048 search1024 search512
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for step in "$@"; do
  case $step in
    tests)  timeout 600 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread --deselect tests/test_tf32_variant.py -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_pytest.log; tail -15 $OUT/${TAG}_pytest.log ;;
    tf32)   timeout 240 python -m pytest tests/test_tf32_variant.py -m gpu -q -s --timeout 200 --timeout-method thread -p no:cacheprovider > $OUT/${TAG}_tf32.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_tf32.log; tail -12 $OUT/${TAG}_tf32.log | cut -c1-1500 ;;
    smoke)  timeout 200 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_smoke.log; tail -4 $OUT/${TAG}_smoke.log ;;
    bench)  timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err; head -c 1500 $OUT/${TAG}_bench.json ;;
    ref)    timeout 300 python bench.py --impl reference > $OUT/${TAG}_ref.json 2> $OUT/${TAG}_ref.err; echo "ref rc=$?"; head -c 600 $OUT/${TAG}_ref.json ;;
    search2048) SEARCH_CHECKPOINT=$OUT/${TAG}_plan2048.json timeout $(( ${SEARCH_S:-420} + 240 )) python tools/sched_search.py tools/sched_exp/libp6d_unsched.so adds_cta_kernelILi256ELi8ELi2ELi0ELi0EE 2 2048 65536 ${SEARCH_S:-420} 6d-pose-estimation_b200/csrc/sched_plan_n2048.json > $OUT/${TAG}_search2048.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_search2048.log; tail -4 $OUT/${TAG}_search2048.log | cut -c1-600 ;;
    search1024) SEARCH_CHECKPOINT=$OUT/${TAG}_plan1024.json timeout $(( ${SEARCH_S:-240} + 240 )) python tools/sched_search.py tools/sched_exp/libp6d_unsched.so adds_cta_kernelILi256ELi4ELi4ELi0ELi0EE 1 1000 262144 ${SEARCH_S:-240} 6d-pose-estimation_b200/csrc/sched_plan_n1024.json 3,0 > $OUT/${TAG}_search1024.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_search1024.log; tail -4 $OUT/${TAG}_search1024.log | cut -c1-600 ;;
    search512)  SEARCH_CHECKPOINT=$OUT/${TAG}_plan512.json timeout $(( ${SEARCH_S:-240} + 240 )) python tools/sched_search.py tools/sched_exp/libp6d_unsched.so adds_cta_kernelILi128ELi4ELi8ELi0ELi0EE 0 500 1048576 ${SEARCH_S:-240} 6d-pose-estimation_b200/csrc/sched_plan_n512.json 8,0 > $OUT/${TAG}_search512.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_search512.log; tail -4 $OUT/${TAG}_search512.log | cut -c1-600 ;;
    ncu)    # launch list + full capture of the headline kernel + the secondary kernels (only after bench exited 0)
            BARGS="--no-sweep --no-pruned --no-secondary --no-cpu-baseline --no-microbench"
            timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 $BARGS > $OUT/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
            timeout 400 ncu --set full --clock-control none --import-source on -k regex:adds_cta --launch-skip 3 -c 1 -f -o $OUT/adds_${TAG} python bench.py --steps 2 --warmup 3 $BARGS > $OUT/${TAG}_ncu_adds.log 2>&1; echo "adds capture rc=$?"
            timeout 600 ncu --set full --clock-control none --import-source on -k regex:'add_pose|pose_loss|pinhole|depth|detection|add_backward|quat|tf32|synth|adds_cta|adds_pruned' -f -o $OUT/secondary_${TAG} python tools/profile_small.py > $OUT/${TAG}_ncu_secondary.log 2>&1; echo "secondary capture rc=$?"; tail -2 $OUT/${TAG}_ncu_secondary.log ;;
  esac
done
