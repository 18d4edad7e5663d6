"""Measured denominators and latencies (ccqp_fp64_peak / ccqp_microbench): prints one JSON object.
    python tools/microbench.py > gpurun_out/<tag>_microbench.json"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ccqppy_b200 import _capi

h = _capi.Handle()
out = dict(fp64_peak_TFLOPs={"%dx%d" % (b, t): h.fp64_peak(b, t) for b, t in ((8, 256), (4, 256), (2, 256), (1, 256), (6, 64), (8, 64),
                                                                           (12, 64), (16, 64), (1, 32), (4, 32))},
           cycles_per_dependent_op=h.microbench(), sm_count=h.info()["sm_count"],
           note="fp64_peak: threads per SM = blocks x threads, 8 independent DFMA chains each; probes: one warp, clock64")
print(json.dumps(out, indent=1))
