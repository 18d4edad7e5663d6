#!/bin/bash
out=gpurun_out; tag=r02p
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_emulated_ranks.py -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/${tag}_pytest.log
