#!/bin/bash
# round 2, session 3, pass e: declared-symmetric upload + the whole bench line
out=gpurun_out; tag=r03e
timeout 600 python -m pytest tests/test_gpu_upload.py -q -x > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03e_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print(d["value"], "e2e", e["value"], "one at a time", e["one_at_a_time"]["value"], "declared", e["declared_symmetric"], "launches", d["gpu_launches"])
print({k: (round(v["qps"] / 1e6, 2), round(v["frac"], 3)) for k, v in d["batched"].items() if isinstance(v, dict) and "qps" in v})
print(d.get("parity"), d["roofline"])
PY
