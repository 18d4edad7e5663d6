"""Read an .ncu-rep here (no GPU): per-opcode totals and the execution-count structure of one kernel.
    python tools/ncu_src.py gpurun_out/x.ncu-rep <kernel regex> [launch index]"""
import csv, io, subprocess, sys
from collections import Counter, defaultdict
rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1])
hdr, data = rows[1], [r for r in rows[2:] if len(r) == len(rows[1])]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
def opname(src, full=False):
    p = src.split()
    if not p: return "?"
    o = p[1] if p[0].startswith("@") and len(p) > 1 else p[0]
    return ".".join(o.split(".")[:3]) if full else o.split(".")[0]
byop = defaultdict(lambda: [0, 0, 0])
for r in data:
    b = byop[opname(r[ix["Source"]], True)]
    b[0] += f(r, "L1 Wavefronts Shared"); b[1] += f(r, "# Samples"); b[2] += f(r, "Instructions Executed")
tot = sum(b[1] for b in byop.values())
print("total samples %d, warp instructions %d" % (tot, sum(b[2] for b in byop.values())))
for op, b in sorted(byop.items(), key=lambda kv: -kv[1][1])[:18]:
    print("%-22s samples %5.1f%%  insts %11.0f  smem wavefronts/inst %.2f" % (op, 100 * b[1] / tot, b[2], b[0] / max(b[2], 1)))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = {h: sum(f(r, h) for r in data) for h in stalls}
print("stalls:", " ".join("%s=%.1f%%" % (k[6:], 100 * v / max(sum(st.values()), 1)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
print("---- structure (runs of equal execution count)")
prev, run = None, []
def flush():
    if run:
        c = Counter(x[0] for x in run)
        print("%10s x%-4d samples=%-6d %s" % (prev, len(run), sum(x[1] for x in run), " ".join("%s:%d" % kv for kv in c.most_common(10))))
for r in data:
    ex = r[ix["Instructions Executed"]]
    if ex != prev:
        flush(); run = []; prev = ex
    run.append((opname(r[ix["Source"]]), int(f(r, "# Samples"))))
flush()
