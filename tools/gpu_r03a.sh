#!/bin/bash
# round 2, session 3, pass a: the symmetric one-warp-per-problem batched kernels (csrc/batched_sym.cu)
out=gpurun_out; tag=r03a
timeout 600 python -m pytest tests/test_gpu_batched.py -q -x -k "symmetric" > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/${tag}_pytest.log
timeout 300 python tools/bench_batched.py > $out/${tag}_batched.json 2> $out/${tag}_batched.err; echo "batched rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r03a_batched.json"))
for k, v in d.items():
    if isinstance(v, dict) and "qps" in v: print(k, round(v["qps"] / 1e6, 2), "M QP/s", round(v["ms"], 3), "ms mv", round(v["mean_mv"], 2), v["bound"], round(v["frac"], 3), v["converged"])
PY
for c in 1 2 4 6 8; do echo "ctas/SM $c"; CCQP_BATCHED_CTAS_PER_SM=$c timeout 200 python tools/bench_batched.py 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
print({k: round(v['qps']/1e6,2) for k,v in d.items() if isinstance(v,dict) and 'qps' in v})"; done
