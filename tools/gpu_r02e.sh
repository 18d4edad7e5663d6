#!/bin/bash
# round 2, pass e: batched tests (incl. the shared-table entry), occupancy sweep of the batched kernel
out=gpurun_out; tag=r02e
mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_batched.py -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/${tag}_pytest.log
for c in 1 2 3 4 5 6; do
  echo "CTAS_PER_SM=$c"; CCQP_BATCHED_CTAS_PER_SM=$c timeout 300 python tools/bench_batched.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:(round(v['qps']/1e6,2), round(v['ms'],3), round(v['mean_mv'],2)) for k,v in d.items() if isinstance(v,dict) and 'qps' in v})"
done
