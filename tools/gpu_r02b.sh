#!/bin/bash
# round 2, second GPU pass: emulated ranks, new CSR phase, sparse bench
out=gpurun_out; tag=r02b
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_emulated_ranks.py tests/test_gpu_sparse.py tests/test_gpu_errors.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $out/${tag}_pytest.log
timeout 600 python tools/bench_sparse.py > $out/${tag}_sparse.json 2> $out/${tag}_sparse.err; echo "sparse rc=$?"; cat $out/${tag}_sparse.json
for g in 1 2 4 8 16; do echo "CSR_GROUP=$g"; CCQP_CSR_GROUP=$g timeout 300 python - <<'PY' 2>/dev/null
import json, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, bench
r = bench.bench_sparse(torch.device("cuda", 0), cases=((1 << 20, 24, 8),), solvers_=("SPG",))
for k, v in r.items(): print(k, {a: (round(b["GBps"]), round(b["us_per_matvec"], 1)) for a, b in v.items() if isinstance(b, dict)})
PY
done
