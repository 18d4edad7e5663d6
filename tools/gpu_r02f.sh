#!/bin/bash
# round 2, pass f: software-pipelined CSR phase; DMMA probes; L2-resident slice of A
out=gpurun_out; tag=r02f
mkdir -p $out
python tools/microbench.py > $out/${tag}_microbench.json 2> $out/${tag}_microbench.err; echo "microbench rc=$?"; python -c "
import json; d=json.load(open('$out/${tag}_microbench.json'))['cycles_per_dependent_op']; print({k:round(v,1) for k,v in d.items()})"
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_emulated_ranks.py tests/test_gpu_parity.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log
CCQP_DEBUG_TIMING=1 timeout 300 python tools/profile_csr.py --solve 2>&1 | grep -v Warn | tail -5
for g in 1 4 8; do echo "CSR_GROUP=$g"; CCQP_CSR_GROUP=$g timeout 300 python tools/profile_csr.py 2>&1 | grep "csr gemv"; done
echo "L1 off"; CCQP_CSR_L1=0 timeout 300 python tools/profile_csr.py 2>&1 | grep "csr gemv"
timeout 600 python tools/bench_sparse.py > $out/${tag}_sparse.json 2> $out/${tag}_sparse.err; echo "sparse rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02f_sparse.json"))
for k, v in d.items(): print(k, {a: (round(b["GBps"]), round(b["us_per_matvec"], 1), b["mv"]) for a, b in v.items() if isinstance(b, dict)})
PY
for mb in 0 32 48 64 80 96; do echo "L2_RESIDENT_MB=$mb"; CCQP_L2_RESIDENT_MB=$mb timeout 300 python tools/bench_n4096.py 2>/dev/null; done
for mb in 0 64; do echo "n=6144 L2_RESIDENT_MB=$mb"; CCQP_L2_RESIDENT_MB=$mb timeout 300 python tools/bench_n4096.py 6144 2>/dev/null; done
for mb in 0 64 96; do echo "bench L2_RESIDENT_MB=$mb"; CCQP_L2_RESIDENT_MB=$mb python bench.py --no-batched --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['achieved'], d.get('apgd',{}).get('GBps_per_gpu'))"; done
timeout 900 python -m pytest tests/test_gpu_batched.py -x -q > $out/${tag}_pytest_batched.log 2>&1; echo "pytest batched rc=$?"; tail -3 $out/${tag}_pytest_batched.log
echo "batched (shuffle sums)"; timeout 300 python tools/bench_batched.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:(round(v['qps']/1e6,2), round(v['ms'],3)) for k,v in d.items() if isinstance(v,dict) and 'qps' in v})"
echo "batched (DMMA sums)"; CCQP_B200_LIB=$PWD/ccqppy_b200/csrc/variants/libccqp_dmma.so timeout 300 python tools/bench_batched.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:(round(v['qps']/1e6,2), round(v['ms'],3)) for k,v in d.items() if isinstance(v,dict) and 'qps' in v})"
CCQP_B200_LIB=$PWD/ccqppy_b200/csrc/variants/libccqp_dmma.so timeout 900 python -m pytest tests/test_gpu_batched.py -x -q > $out/${tag}_pytest_batched_dmma.log 2>&1; echo "pytest batched dmma rc=$?"; tail -3 $out/${tag}_pytest_batched_dmma.log
REPS=1 python tools/profile_csr.py > $out/${tag}_csr_plain.log 2>&1 &&
REPS=1 ncu --set full --clock-control none --import-source on -k regex:'dense_kernel' -s 1 -c 1 -f -o $out/${tag}_csr python tools/profile_csr.py > $out/${tag}_csr_ncu.log 2>&1
echo "ncu rc=$?"
