#!/bin/bash
# round 2, pass h: whole GPU suite (n <= 128 batched layout, CSR many-warps build), phase timing at n = 4096
out=gpurun_out; tag=r02h
mkdir -p $out
timeout 1700 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $out/${tag}_pytest.log
timeout 300 python tools/profile_csr.py 2>&1 | grep -v Warn | tail -3
echo "n=128 batched"; timeout 300 python - <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, bench
for n, B in ((128, 16384), (96, 16384), (64, 65536)):
    bench.NB, bench.BATCH = n, B
    d = bench.bench_batched(torch.device("cuda", 0), steps=3, warmup=2)
    print(n, {k: (round(v["qps"] / 1e6, 3), round(v["ms"], 3), round(v["mean_mv"], 1), round(v.get("fp64_frac", 0), 3)) for k, v in d.items() if isinstance(v, dict) and "qps" in v})
PY
echo "phase timing n=4096 SPG"; CCQP_DEBUG_TIMING=1 timeout 300 python tools/bench_n4096.py 2>&1 | grep "ccqp timing" | tail -8
for g in 1 2 4 8 16; do echo "CSR_GROUP=$g"; CCQP_CSR_GROUP=$g timeout 300 python tools/profile_csr.py 2>&1 | grep "csr gemv\|max rel"; done
CCQP_DEBUG_TIMING=1 timeout 300 python tools/profile_csr.py --solve 2>&1 | grep -v Warn | tail -3
REPS=1 python tools/profile_csr.py > $out/${tag}_csr_plain.log 2>&1 &&
REPS=1 ncu --set full --clock-control none --import-source on -k regex:'dense_kernel' -s 1 -c 1 -f -o $out/${tag}_csr python tools/profile_csr.py > $out/${tag}_csr_ncu.log 2>&1
echo "ncu rc=$?"
