#!/bin/bash
# round 2, multi-GPU pass (gpurun --gpus N): sharded parity on the final code, bench at N (own arm + reference arm)
# usage: bash tools/gpu_r02_multi.sh <N> [tag]
N=${1:-2}; tag=${2:-r02mg$N}; out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.csv 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > $out/${tag}_check.log 2> $out/${tag}_check.err; echo "multi_gpu_check rc=$?"; grep -a "MULTI_GPU_CHECK" $out/${tag}_check.log | tail -2
grep -ac " ok " $out/${tag}_check.log; grep -a "FAIL" $out/${tag}_check.log | head -5
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_errors.py -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; python - <<PY
import json
d = json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step")}, d["roofline"]["frac"], d.get("parity", {}).get("ok"), d.get("apgd", {}).get("value"), {k: round(v["qps"]/1e6, 1) for k, v in d.get("batched", {}).items() if isinstance(v, dict) and "qps" in v})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"; cat $out/${tag}_bench_ref.json | cut -c1-600
if [ "$N" = "8" ]; then
  python bench.py --gpus 1 --no-batched --no-sparse --no-cpu-baseline --steps 6 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench n1 rc=$?"; python -c "
import json; d=json.loads(open('$out/${tag}_bench_n1.json').read().strip().splitlines()[-1]); print('N=1 same box:', d['value'], d['roofline']['frac'])"
fi
