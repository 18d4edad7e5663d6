#!/bin/bash
# round 2, pass u: MPRGP on CSR with the phase outlined + register copy of the mbarrier parities
out=gpurun_out; tag=r02u
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_emulated_ranks.py -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log
CCQP_DEBUG_TIMING=1 timeout 300 python tools/mprgp_sparse_timing.py 2>&1 | grep -v Warn | tail -3
timeout 120 python tools/profile_csr.py 2>&1 | grep "csr gemv\|max rel"
timeout 600 python tools/bench_sparse.py > $out/${tag}_sparse.json 2> $out/${tag}_sparse.err; echo "sparse rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02u_sparse.json"))
for k, v in d.items(): print(k, {a: (round(b["GBps"]), round(b["us_per_matvec"], 1), b["mv"]) for a, b in v.items() if isinstance(b, dict)})
PY
