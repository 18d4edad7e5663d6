#!/bin/bash
# round 2, first GPU pass: tests, smoke, microbench, bench (own + reference arm), batched ncu with source
out=gpurun_out; tag=r02a
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.csv 2>&1
python tools/microbench.py > $out/${tag}_microbench.json 2> $out/${tag}_microbench.err; echo "microbench rc=$?"; cat $out/${tag}_microbench.json
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $out/${tag}_pytest_gpu.log
python __graft_entry__.py --smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cat $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"; cat $out/${tag}_bench_ref.json
python tools/profile_target.py batched > $out/${tag}_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'batched_kernel' -c 2 -f -o $out/${tag}_batched python tools/profile_target.py batched > $out/${tag}_prof_ncu.log 2>&1
echo "ncu rc=$?"; cat $out/${tag}_prof_plain.log
