#!/bin/bash
# round 2, session 3, last pass: upload fixes (declared symmetry at every size, pending mirror reset) -> tests, DRAM traffic of the
# dense kernel for the new source hash, bench line
out=gpurun_out; tag=${1:-r03x}
timeout 600 python -m pytest tests/test_gpu_upload.py tests/test_gpu_sparse.py -q -x > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/${tag}_pytest.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'dense_kernel' -s 3 -c 1 -f -o $out/${tag}_dense_spg \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-batched --no-sparse > $out/${tag}_ncu_dense.log 2>&1
echo "ncu dense rc=$?"
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-200 $out/${tag}_bench.json
