#!/bin/bash
# round 2, pass g: batched kernel (2-deep queue, staged vectors, DMMA sums variant), grid-sync closer, probes
out=gpurun_out; tag=r02g
mkdir -p $out
python tools/microbench.py > $out/${tag}_microbench.json 2> $out/${tag}_microbench.err; echo "microbench rc=$?"; python -c "
import json; d=json.load(open('$out/${tag}_microbench.json'))['cycles_per_dependent_op']; print({k:round(v,1) for k,v in d.items()})"
timeout 900 python -m pytest tests/test_gpu_batched.py tests/test_gpu_parity.py tests/test_gpu_emulated_ranks.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log
pr='import json,sys
d=json.loads(sys.stdin.read()); print({k:(round(v["qps"]/1e6,2), round(v["ms"],3)) for k,v in d.items() if isinstance(v,dict) and "qps" in v})'
echo "batched (shuffle sums)"; timeout 300 python tools/bench_batched.py 2>/dev/null | python -c "$pr"
echo "batched (DMMA sums)"; CCQP_B200_LIB=$PWD/ccqppy_b200/csrc/variants/libccqp_dmma.so timeout 300 python tools/bench_batched.py 2>/dev/null | python -c "$pr"
CCQP_B200_LIB=$PWD/ccqppy_b200/csrc/variants/libccqp_dmma.so timeout 900 python -m pytest tests/test_gpu_batched.py -x -q > $out/${tag}_pytest_dmma.log 2>&1; echo "pytest dmma rc=$?"; tail -3 $out/${tag}_pytest_dmma.log
echo "n=4096"; timeout 300 python tools/bench_n4096.py 2>/dev/null
echo "bench"; python bench.py --no-batched --no-sparse --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['achieved'], d.get('apgd',{}).get('GBps_per_gpu'), d['parity']['ok'])"
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_errors.py -x -q > $out/${tag}_pytest_sparse.log 2>&1; echo "pytest sparse rc=$?"; tail -3 $out/${tag}_pytest_sparse.log
CCQP_DEBUG_TIMING=1 timeout 300 python tools/profile_csr.py --solve 2>&1 | grep -v Warn | tail -6
for g in 2 4 8 16 32; do echo "CSR_GROUP=$g"; CCQP_CSR_GROUP=$g timeout 300 python tools/profile_csr.py 2>&1 | grep "csr gemv"; done
timeout 600 python tools/bench_sparse.py > $out/${tag}_sparse.json 2> $out/${tag}_sparse.err; echo "sparse rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02g_sparse.json"))
for k, v in d.items(): print(k, {a: (round(b["GBps"]), round(b["us_per_matvec"], 1), b["mv"]) for a, b in v.items() if isinstance(b, dict)})
PY
REPS=1 python tools/profile_csr.py > $out/${tag}_csr_plain.log 2>&1 &&
REPS=1 ncu --set full --clock-control none --import-source on -k regex:'dense_kernel' -s 1 -c 1 -f -o $out/${tag}_csr python tools/profile_csr.py > $out/${tag}_csr_ncu.log 2>&1
echo "ncu rc=$?"
