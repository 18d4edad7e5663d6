"""profiles/dram_traffic.json from an ncu summary (tools/summarize_ncu.py full ...) of ONE dense_kernel<SPG> launch of the bench
workload, stamped with the hash of the CUDA sources of THIS tree (bench.kernel_source_hash):
    python tools/update_dram_traffic.py profiles/<tag>_ncu_dense_spg_n32768.csv <gpurun call tag>"""
import csv
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

src, tag = sys.argv[1], sys.argv[2]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
vals = {}
for row in csv.reader(open(src)):
    if len(row) >= 3 and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        vals[row[0]] = float(row[2]) * scale[row[1]]
n = 32768
mv = 56
out = dict(dense_spg_bytes_per_launch=vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"], n=n,
           kernel_source_sha256=bench.kernel_source_hash(),
           source="%s: dram__bytes_read.sum + dram__bytes_write.sum of ONE dense_kernel<SPG> launch of the bench workload (%d mat-vecs, "
                  "n=%d) under ncu --set full --clock-control none (gpurun call %s); 64 MB of A stay in L2 between mat-vecs, hence "
                  "slightly below the algorithmic bytes" % (os.path.relpath(src, ROOT), mv, n, tag),
           algorithmic_bytes_per_launch=mv * (8.0 * n * n + 16.0 * n))
json.dump(out, open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
print(out)
