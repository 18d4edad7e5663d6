"""Config 4 only (65536 box-QPs of n=64, BBPGD and SPG): prints the `batched` object of bench.py."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
print(json.dumps(bench.bench_batched(torch.device("cuda", 0), steps=5, warmup=3)))
