#!/bin/bash
# round 2, session 3, pass d: symmetric upload (csrc/upload.cu): tests, speed of the host check on the box's cores, bench line
out=gpurun_out; tag=r03d
nproc; lscpu | grep "Model name"
g++ -O3 -std=c++17 -pthread -I ccqppy_b200/csrc tools/symcheck_time.cpp -o /tmp/symcheck_time && for t in 1 8 16 0; do /tmp/symcheck_time 32768 $t | tail -1; done
timeout 600 python -m pytest tests/test_gpu_upload.py -q -x > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log
timeout 900 python bench.py --no-batched > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03d_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"], d.get("parity"), d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
CCQP_SYM_UPLOAD=0 timeout 900 python bench.py --no-batched --no-cpu-baseline > $out/${tag}_bench_full_upload.json 2> $out/${tag}_bench_full_upload.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03d_bench_full_upload.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"])
PY
