#!/bin/bash
# round 2, session 3, pass b: symmetric batched kernels after the register-pressure fix (b / lb / ub in shared memory, 4 DFMA chains)
out=gpurun_out; tag=r03b
timeout 600 python -m pytest tests/test_gpu_batched.py -q -x -k "symmetric" > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/${tag}_pytest.log
for c in 1 8; do echo "ctas/SM $c"; CCQP_BATCHED_CTAS_PER_SM=$c timeout 200 python tools/bench_batched.py 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
print({k: round(v['qps']/1e6,2) for k,v in d.items() if isinstance(v,dict) and 'qps' in v})"; done
