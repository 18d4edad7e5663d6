"""Dump GPU-vs-golden for every golden solver case (diagnostic; run under gpurun)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as pr
from helpers import run_gpu
from test_oracle_golden import META, SOL, case_inputs
rows = []
for c in META:
    A, b, tab, x0 = case_inputs(c)
    out = run_gpu(c["solver"], A, b, tab, x0=x0, tol=c["tol"], max_mv=c["max_mv"], step=c["step"], spg_seed=c["spg_seed"])
    gold = SOL[c["name"]]
    err = float(np.linalg.norm(out["solution"] - gold) / max(np.linalg.norm(gold), 1e-300))
    rows.append(dict(name=c["name"], mv=out["mv"], mv_gold=c["mv"], conv=out["converged"], conv_gold=c["converged"],
                     err=err, res=out["residual"], res_gold=float.fromhex(c["residual"])))
    flag = "" if (out["mv"] == c["mv"] and err < 1e-9) else "   <<<"
    print("%-40s mv %5d / %5d  err %.2e res %.3e / %.3e%s" % (c["name"], out["mv"], c["mv"], err, out["residual"], rows[-1]["res_gold"], flag))
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "diag_parity.json"), "w"))
