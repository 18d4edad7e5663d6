#!/bin/bash
# round 2, session 3: final single-GPU evidence pass for the final sources (tests, smoke, bench own + reference arm, launch list,
# ncu --set full of ONE dense_kernel<SPG> launch for dram_traffic.json and of the symmetric batched kernel)
out=gpurun_out; tag=${1:-r03z}
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.csv 2>&1
nproc > $out/${tag}_host.txt; lscpu | grep "Model name" >> $out/${tag}_host.txt
timeout 1200 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log
python __graft_entry__.py --smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-300 $out/${tag}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"; cut -c1-200 $out/${tag}_bench_ref.json
# launch list of the bench command itself (cold-cache, serialised: compare shares)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
echo "launches rc=$?"
# one dense_kernel<SPG> launch of the bench workload (n = 32768, 56 mat-vecs) under ncu --set full: DRAM traffic per launch
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'dense_kernel' -s 3 -c 1 -f -o $out/${tag}_dense_spg \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-batched --no-sparse > $out/${tag}_ncu_dense.log 2>&1
echo "ncu dense rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'batched_sym_kernel' -c 2 -f -o $out/${tag}_batched_sym python tools/profile_target.py batched_sym > $out/${tag}_ncu_batched_sym.log 2>&1
echo "ncu batched_sym rc=$?"
