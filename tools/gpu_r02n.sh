#!/bin/bash
out=gpurun_out; tag=r02n
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_emulated_ranks.py -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log
for g in 2 4 8; do echo "CSR_GROUP=$g"; CCQP_CSR_GROUP=$g timeout 300 python tools/profile_csr.py 2>&1 | grep "csr gemv\|max rel"; done
CCQP_DEBUG_TIMING=1 timeout 300 python tools/profile_csr.py --solve 2>&1 | grep -v Warn | tail -2
timeout 600 python tools/bench_sparse.py > $out/${tag}_sparse.json 2> $out/${tag}_sparse.err; echo "sparse rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02n_sparse.json"))
for k, v in d.items(): print(k, {a: (round(b["GBps"]), round(b["us_per_matvec"], 1), b["mv"]) for a, b in v.items() if isinstance(b, dict)})
PY
REPS=1 python tools/profile_csr.py > $out/${tag}_csr_plain.log 2>&1 &&
REPS=1 ncu --set full --clock-control none --import-source on -k regex:'dense_kernel' -s 1 -c 1 -f -o $out/${tag}_csr python tools/profile_csr.py > $out/${tag}_csr_ncu.log 2>&1
echo "ncu rc=$?"
