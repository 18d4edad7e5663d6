#!/bin/bash
# round 2, final single-GPU evidence pass: tests, smoke, bench (own + reference arm), launch list, ncu captures, other configs
out=gpurun_out; tag=${1:-r02z}
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.csv 2>&1
python tools/microbench.py > $out/${tag}_microbench.json 2> $out/${tag}_microbench.err; echo "microbench rc=$?"
timeout 1700 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log
python __graft_entry__.py --smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-400 $out/${tag}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $out/${tag}_bench_ref.json
timeout 900 python tools/bench_configs.py > $out/${tag}_configs.json 2> $out/${tag}_configs.err; echo "configs rc=$?"
timeout 600 python tools/bench_sparse.py > $out/${tag}_sparse.json 2> $out/${tag}_sparse.err; echo "sparse rc=$?"
timeout 300 python tools/bench_n4096.py > $out/${tag}_n4096.json 2>/dev/null; cat $out/${tag}_n4096.json
# launch list of the bench command itself (cold-cache, serialised: compare shares)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_short.json 2> $out/${tag}_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
echo "launches rc=$?"
# one dense_kernel<SPG> launch of the bench workload (n = 32768, 56 mat-vecs) under ncu --set full: DRAM traffic per launch
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-batched --no-sparse > $out/${tag}_plain_dense.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'dense_kernel' -s 3 -c 1 -f -o $out/${tag}_dense_spg \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-batched --no-sparse > $out/${tag}_ncu_dense.log 2>&1
echo "ncu dense rc=$?"
python tools/profile_target.py batched > $out/${tag}_plain_batched.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'batched_kernel' -c 2 -f -o $out/${tag}_batched python tools/profile_target.py batched > $out/${tag}_ncu_batched.log 2>&1
echo "ncu batched rc=$?"
REPS=1 python tools/profile_csr.py > $out/${tag}_plain_csr.log 2>&1 &&
REPS=1 ncu --set full --clock-control none --import-source on -k regex:'dense_kernel' -s 1 -c 1 -f -o $out/${tag}_csr python tools/profile_csr.py > $out/${tag}_ncu_csr.log 2>&1
echo "ncu csr rc=$?"; cat $out/${tag}_plain_csr.log | grep "csr gemv"
for mb in 0 64; do echo "shard mat-vec, L2_RESIDENT_MB=$mb"; CCQP_L2_RESIDENT_MB=$mb timeout 300 python tools/shard_gemv_time.py 2>&1 | grep "P="; done
CCQP_DEBUG_TIMING=1 timeout 300 python tools/profile_csr.py --solve 2>&1 | grep -v Warn | tail -2
