#!/bin/bash
out=gpurun_out; tag=r02d
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_emulated_ranks.py::test_emulated_ranks_csr "tests/test_gpu_parity.py::test_projected_gradient_matches_reference_goldens_and_behaviours" tests/test_gpu_errors.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/${tag}_pytest.log
CCQP_DEBUG_TIMING=1 timeout 300 python tools/profile_csr.py --solve 2>&1 | grep -v Warn | tail -6
timeout 600 python tools/bench_sparse.py > $out/${tag}_sparse.json 2> $out/${tag}_sparse.err; echo "sparse rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02d_sparse.json"))
for k, v in d.items(): print(k, {a: (round(b["GBps"]), round(b["us_per_matvec"], 1), b["mv"]) for a, b in v.items() if isinstance(b, dict)})
PY
for g in 1 4; do echo "CSR_GROUP=$g"; CCQP_CSR_GROUP=$g timeout 300 python tools/profile_csr.py 2>&1 | grep "csr gemv"; done
echo "L1 off"; CCQP_CSR_L1=0 timeout 300 python tools/profile_csr.py 2>&1 | grep "csr gemv"
REPS=1 python tools/profile_csr.py > $out/${tag}_csr_plain.log 2>&1 &&
REPS=1 ncu --set full --clock-control none --import-source on -k regex:'dense_kernel' -s 1 -c 1 -f -o $out/${tag}_csr python tools/profile_csr.py > $out/${tag}_csr_ncu.log 2>&1
echo "ncu rc=$?"
