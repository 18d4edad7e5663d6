#!/bin/bash
out=gpurun_out; tag=r02c
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_emulated_ranks.py::test_emulated_ranks_csr tests/test_ccqppy_shim.py "tests/test_gpu_parity.py::test_benchmark_driver_with_the_gpu_generator" -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/${tag}_pytest.log
timeout 600 python tools/bench_sparse.py > $out/${tag}_sparse.json 2> $out/${tag}_sparse.err; echo "sparse rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c_sparse.json"))
for k, v in d.items(): print(k, {a: (round(b["GBps"]), round(b["us_per_matvec"], 1), b["mv"]) for a, b in v.items() if isinstance(b, dict)})
PY
for g in 1 2 4 8; do echo "CSR_GROUP=$g"; CCQP_CSR_GROUP=$g timeout 300 python - <<'PY' 2>/dev/null
import json, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, bench
r = bench.bench_sparse(torch.device("cuda", 0), cases=((1 << 20, 24, 8),), solvers_=("SPG",))
for k, v in r.items(): print(k, {a: (round(b["GBps"]), round(b["us_per_matvec"], 1)) for a, b in v.items() if isinstance(b, dict)})
PY
done
echo "TMA off"; CCQP_CSR_TMA=0 timeout 300 python - <<'PY' 2>/dev/null
import json, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, bench
r = bench.bench_sparse(torch.device("cuda", 0), cases=((1 << 20, 24, 8),), solvers_=("SPG",))
for k, v in r.items(): print(k, {a: (round(b["GBps"]), round(b["us_per_matvec"], 1)) for a, b in v.items() if isinstance(b, dict)})
PY
