#!/bin/bash
# One GPU validation pass (run under gpurun from the repo root):
#   tools/gpu_round.sh [tag] [steps...]      steps: tests smoke bench ref launches full
# Everything lands in gpurun_out/<tag>_*.
tag=${1:-r01}; shift
steps=${*:-tests smoke bench ref launches full}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.csv 2>&1
for s in $steps; do
  case $s in
    tests)
      python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
      tail -5 $out/${tag}_pytest_gpu.log ;;
    smoke)
      python __graft_entry__.py --smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/${tag}_smoke.log
      tail -4 $out/${tag}_smoke.log ;;
    bench)
      python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
      cat $out/${tag}_bench.json ;;
    ref)
      python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"
      cat $out/${tag}_bench_ref.json ;;
    launches)
      # launch list of the bench command itself (cold-cache, serialised: compare shares)
      python bench.py --steps 2 --warmup 3 > $out/${tag}_bench_short.json 2> $out/${tag}_bench_short.err &&
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
          python bench.py --steps 2 --warmup 3 > $out/${tag}_ncu_launches.log 2>&1
      echo "launches rc=$?" ;;
    full)
      python tools/profile_target.py > $out/${tag}_prof_plain.log 2>&1 &&
      ncu --set full --clock-control none --import-source on -k regex:'dense_kernel|batched_kernel' -c 8 \
          -f -o $out/${tag}_prof python tools/profile_target.py > $out/${tag}_prof_ncu.log 2>&1
      echo "full rc=$?"; cat $out/${tag}_prof_plain.log ;;
  esac
done
