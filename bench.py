#!/usr/bin/env python
"""Headline benchmark of the CCQP projected-gradient hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference arm)

Workload (BASELINE.json config 3): dense SPG-QP, n = 32768, fp64,
A = G G^T / n + I (seed 0, symmetrised to the last bit), b = -A x*, x* = 1 - 4 U, Box[-1,1]^n, tol 1e-5, max 2000 mat-vecs,
SPG uniforms = RandomState(0).  One "step" = one whole solve (about 56 mat-vecs, each streaming
the 8.59 GB Hessian once).  metric = SPG iterations (= mat-vecs executed) per second.

  value : device-resident A (already in HBM), timed with CUDA events around the K solves
  e2e   : the same solves through the public API from pinned HOST buffers: the timed region
          includes the host->device copy of A, b, uniforms and the device->host copy of x
          (`value` of e2e = a pipelined STREAM of solves, `one_at_a_time` = the plain loop).  ccqp_set_matrix finds out
          on the host, while the copy engine works, that this A is symmetric and moves only its upper block triangle
          (`matrix_upload`, `h2d_bytes_per_step` = what actually crossed PCIe; the device copy is bit-identical to a full
          upload); `declared_symmetric` = the same stream for a caller that declares the symmetry (no host-side test)
  roofline : algorithmic bytes of the solver kernel (mat-vecs x (8 n^2 + 16 n)) / its duration,
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  parity : the solution / mat-vec count / converged flag of the timed solve against the CPU port of
           the reference (oracle/) run for the WHOLE solve on the same arrays (N = 1), or against the
           single-GPU kernel's answer on the same problem (N > 1); same for `apgd`
  apgd  : config 3's second solver (CCQPSolverAPGD) on the same problem
  cpu_baseline : the NumPy/OpenBLAS port of the reference (oracle/) on this host's cores, the whole solve
  batched : config 4 (65536 box-QPs of n = 64, BBPGD and SPG) in the persistent per-CTA kernel, each
            against max(HBM bytes / measured HBM peak, mat-vec flops / MEASURED fp64 peak); `*_sym` = the same problems
            through ccqp_solve_batched_sym (declared symmetric Hessians, one warp per problem)
  sparse : operator-form (CSR) Hessian, n = 2^20, ~57 stored entries per row (row f-3)

With N > 1 (torchrun, one rank per GPU) A is row-sharded and the SAME problem is solved by all
ranks together (strong scaling); the exchange of the mat-vec input and of the scalar partial sums
happens inside the solver kernel through NVLink peer memory.
"""
import os
import sys


def _reference_arm_env():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is a CPU measurement that must use all
    host threads, so the BLAS thread limits are lifted BEFORE NumPy loads its BLAS (and again with threadpoolctl)."""
    argv = sys.argv
    if "--impl" in argv and argv[argv.index("--impl") + 1:argv.index("--impl") + 2] == ["reference"] or "--impl=reference" in argv:
        for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "GOTO_NUM_THREADS"):
            os.environ.pop(k, None)


_reference_arm_env()

import argparse      # noqa: E402
import hashlib       # noqa: E402
import json          # noqa: E402
import subprocess    # noqa: E402
import threading     # noqa: E402
import time          # noqa: E402

import numpy as np   # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_DENSE = int(os.environ.get("CCQP_BENCH_N", 32768))
TOL, MAX_MV, SEED = 1e-5, 2000, 0
BATCH, NB = int(os.environ.get("CCQP_BENCH_BATCH", 65536)), 64
REF_SAMPLE_MV = 8          # mat-vecs per step of the reference arm


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first
# communicator), so the process's fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved,
# real stdout.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_config(n, world):
    """`config` of the JSON line: byte-identical in this repo's arm and in the reference arm."""
    return dict(workload="dense SPG-QP n=%d fp64, A=GG^T/n+I seed 0, b=-A(1-4U), Box[-1,1], tol 1e-5, max_mv 2000, "
                         "uniforms RandomState(0)" % n,
                n_gpus=world,
                l2="inputs (%.2f GB of A per GPU) far exceed the 126 MB L2; no flush needed" % (8e-9 * (n // world) * n))


def kernel_source_hash():
    """sha256 over the CUDA sources: stamps measurements that were taken under ncu for a particular kernel build."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "ccqppy_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh")):
            h.update(name.encode()); h.update(open(os.path.join(csrc, name), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "ccqp_b200.h"), "rb").read())
    return h.hexdigest()[:16]


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.t0 = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """The timed region starts now (the sampler itself is started before the warm-up: nvidia-smi needs a few hundred
        milliseconds to come up, more than a short timed region at 8 GPUs lasts)."""
        self.t0 = time.time()

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = self.t0 if self.t0 is not None else 0.0
        rows = [r for t, r in self.rows if t >= t0]
        window = "timed region"
        if not rows:        # the timed region was shorter than one sampling period: the warm-up ran the same kernels
            rows, window = [r for _, r in self.rows], "warm-up + timed region (the timed region is shorter than one sampling period)"
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), window=window)


# ---------------------------------------------------------------------------------------------
def make_dense_problem(n, device):
    """Seeded 'shift' problem generated on the GPU in fp64 (the n^3 product is the step BEFORE the
    path; generating it on the host would take minutes)."""
    import torch
    g = torch.Generator(device=device).manual_seed(SEED)
    G = torch.randn((n, n), generator=g, device=device, dtype=torch.float64)
    A = torch.empty((n, n), device=device, dtype=torch.float64)
    torch.matmul(G, G.t(), out=A)
    del G
    A.div_(n)
    T = A.t().contiguous()          # a Hessian is symmetric: make it so to the last bit (a GEMM's G G^T need not be)
    A.add_(T).mul_(0.5)
    del T
    A.diagonal().add_(1.0)
    xs = 1.0 - 4.0 * torch.rand(n, generator=g, device=device, dtype=torch.float64)
    b = -(A @ xs)
    return A, b


def spg_uniform_stream(count):
    return np.random.RandomState(SEED).random_sample(count)


def parity_record(against, x_gpu, mv_gpu, conv_gpu, x_ref, mv_ref, conv_ref):
    """north_star's tolerance: solution within 1e-9 relative, the same converged flag, mat-vec count within 2 %."""
    x_gpu, x_ref = np.asarray(x_gpu, dtype=np.float64), np.asarray(x_ref, dtype=np.float64)
    rel = float(np.linalg.norm(x_gpu - x_ref) / max(np.linalg.norm(x_ref), 1e-300))
    mv_ok = abs(int(mv_gpu) - int(mv_ref)) <= max(1, round(0.02 * int(mv_ref)))
    return dict(against=against, mv_gpu=int(mv_gpu), mv_ref=int(mv_ref), mv_within_2pct=bool(mv_ok), rel_err=rel,
                converged_gpu=bool(conv_gpu), converged_equal=bool(conv_gpu) == bool(conv_ref),
                ok=bool(mv_ok and rel <= 1e-9 and bool(conv_gpu) == bool(conv_ref)), tolerance="rel_err <= 1e-9, mv within 2 %")


def run_reference_arm(args, rank):
    """The reference's own algorithm on the host cores: the NumPy/OpenBLAS port in oracle/ (the
    Python reference cannot travel to the GPU box; the port is bit-identical to it in the build
    container, see oracle/gen_golden.py), `kind: "port"`.  Each step = REF_SAMPLE_MV mat-vecs of the workload."""
    if rank != 0:
        return
    import torch
    from threadpoolctl import threadpool_info, threadpool_limits
    from oracle import ccqp_oracle as orc
    import problems as pr
    n = N_DENSE
    if torch.cuda.is_available():
        A_d, b_d = make_dense_problem(n, "cuda:0")
        A, b = A_d.cpu().numpy(), b_d.cpu().numpy()
        del A_d, b_d
        torch.cuda.empty_cache()
    else:
        A, b = pr.shift_problem(n, SEED)
    tab = pr.box_table(n)
    uni = spg_uniform_stream(MAX_MV)
    avail = host_threads()

    def step():
        o = orc.solve(orc.SPG, A, b, blocks=tab.blocks, params=tab.params, tol=TOL, max_mv=REF_SAMPLE_MV, uniforms=uni)
        return o["gemv"]
    with threadpool_limits(limits=avail, user_api="blas"):
        cores = max([i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"] or [1])
        if cores != avail:
            log("[bench] reference arm: BLAS uses %d threads, the host offers %d" % (cores, avail))
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        mvs = sum(step() for _ in range(args.steps))
        dt = time.perf_counter() - t0
    value = mvs / dt
    sample = "%d steps x %d mat-vecs of the n=%d SPG solve (solver capped with max_mv), NumPy/OpenBLAS port of the reference" % (
        args.steps, REF_SAMPLE_MV, n)
    line = dict(metric="spg_iterations_per_s_dense_n%d_fp64" % n, value=value, unit="iterations/s", impl="reference",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * dt / args.steps,
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=workload_config(n, args.gpus), sample=sample,
                cpu_baseline=dict(value=value, unit="iterations/s", cores=cores, host_threads=avail, kind="port", sample=sample),
                e2e=dict(value=value, unit="iterations/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                effective_GBps=value * (8.0 * n * n + 16.0 * n) / 1e9)
    emit(line)


# ---------------------------------------------------------------------------------------------
_FP64_PEAK = {}


def fp64_peak(device_index):
    """Measured DFMA throughput (ccqp_fp64_peak): SM full, and at the batched kernel's own occupancy (6 CTAs x 64)."""
    if device_index not in _FP64_PEAK:
        from ccqppy_b200 import _capi
        h = _capi.Handle(device_index)
        _FP64_PEAK[device_index] = dict(full=h.fp64_peak(8, 256), at_batched_occupancy=h.fp64_peak(6, 64))
        h.close()
    return _FP64_PEAK[device_index]


def bench_batched(device, steps, warmup, batch=None, seed=1, reduce_max=None):
    """Config 4.  `batch` problems on this GPU; with `reduce_max` (multi-GPU) the slowest rank's time counts and
    the rates are for the whole job of BATCH problems.  Roofline per solver:
    T_roof = max(HBM bytes / measured HBM peak, 2 n^2 x mat-vecs executed / measured fp64 peak); frac = T_roof / T."""
    import torch
    from ccqppy_b200 import solvers
    local = BATCH if batch is None else batch
    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    chunk = 8192
    A = torch.empty((local, NB, NB), device=device, dtype=torch.float64)
    b = torch.empty((local, NB), device=device, dtype=torch.float64)
    for s in range(0, local, chunk):
        e = min(local, s + chunk)
        G = torch.randn((e - s, NB, NB), generator=g, device=device, dtype=torch.float64)
        A[s:e] = G @ G.transpose(1, 2) / NB + torch.eye(NB, device=device, dtype=torch.float64)
        A[s:e] = 0.5 * (A[s:e] + A[s:e].transpose(1, 2))        # exactly symmetric (the `_sym` rows read the upper block triangle only)
        xs = 1 - 4 * torch.rand((e - s, NB), generator=g, device=device, dtype=torch.float64)
        b[s:e] = -(A[s:e] @ xs.unsqueeze(-1)).squeeze(-1)
    lb, ub = -torch.ones_like(b), torch.ones_like(b)
    K = 256
    uni = torch.rand((local, K), generator=g, device=device, dtype=torch.float64)
    peak, _ = measured_peak()
    fpk = fp64_peak(device.index or 0)
    ranks = 1
    # `X`: ccqp_solve_batched (any A, one CTA of 64 threads per problem); `X_sym`: ccqp_solve_batched_sym (the caller declares A
    # symmetric; one warp per problem, upper block triangle only) -- same problems, same tolerance
    for name, cls, sym in (("BBPGD", solvers.CCQPSolverBBPGD, False), ("SPG", solvers.CCQPSolverSPG, False),
                           ("BBPGD_sym", solvers.CCQPSolverBBPGD, True), ("SPG_sym", solvers.CCQPSolverSPG, True)):
        s = cls(1e-8, 5000)
        s.quiet = True
        times = []
        for it in range(warmup + steps):
            s.solve_batched(A, b, lb, ub, uniforms=uni, symmetric=sym)
            if it >= warmup:
                times.append(s.solution_gpu_time)
        t = float(np.mean(times))
        hbm = s.solution_hbm_bytes
        flops = 2.0 * NB * NB * s.solution_gemv_count
        total = local
        if reduce_max is not None:      # slowest rank's time; bytes, flops and problems of all ranks
            t, hbm, flops, total, ranks = reduce_max(t, hbm, flops, local)
        t_hbm, t_f64 = hbm / (peak * 1e9 * ranks), flops / (fpk["full"] * 1e12 * ranks)
        out[name] = dict(qps=total / t, ms=1e3 * t, mean_mv=float(np.mean(s.solution_num_matrix_vector_multiplications)),
                         converged=bool(np.all(s.solution_converged)), hbm_GBps=hbm / t / 1e9, hbm_frac=t_hbm / t,
                         fp64_TFLOPs=flops / t / 1e12, fp64_frac=t_f64 / t,
                         bound="hbm" if t_hbm >= t_f64 else "fp64", frac=max(t_hbm, t_f64) / t,
                         roofline_qps=total / max(t_hbm, t_f64))
    out["fp64_peak_TFLOPs"] = dict(fpk, how="ccqp_fp64_peak: 8 independent DFMA chains per thread, best of 3, CUDA events; "
                                           "`full` = 8 x 256 threads per SM (the roofline denominator), the other at 6 x 64")
    return out


def batched_cpu_baseline(sample=192):
    """The reference's way of doing config 4: a Python loop of solve() calls, one problem at a time (the port in
    oracle/, one core -- the reference has no batching or threading), on a bounded sample of the same generator."""
    from oracle import ccqp_oracle as orc
    import problems as pr
    tab = pr.box_table(NB)
    probs = [pr.shift_problem(NB, 1000 + i) for i in range(sample)]
    out = {}
    for name, sid in (("BBPGD", orc.BBPGD), ("SPG", orc.SPG)):
        t0 = time.perf_counter()
        mv = 0
        for i, (A, b) in enumerate(probs):
            mv += orc.solve(sid, A, b, blocks=tab.blocks, params=tab.params, tol=1e-8, max_mv=5000,
                            uniforms=pr.spg_uniforms(i, 512))["mv"]
        dt = time.perf_counter() - t0
        out[name] = dict(qps=sample / dt, mean_mv=mv / sample)
    out.update(cores=1, kind="port", sample="%d problems of n=%d solved one after the other, as the reference would" % (sample, NB))
    return out


def sparse_matrix(n, band, extra, seed=0):
    """Banded + random symmetric, strictly diagonally dominant CSR matrix (contact-style Hessians are sparse)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    offs = np.arange(1, band + 1)
    diags = [rng.standard_normal(n - o) * 0.5 / band for o in offs]
    B = sp.diags(diags, offs, shape=(n, n), format="csr")
    if extra:
        rows = np.repeat(np.arange(n), extra // 2)
        cols = rng.integers(0, n, rows.size)
        B = B + sp.csr_matrix((rng.standard_normal(rows.size) * 0.5 / band, (rows, cols)), shape=(n, n))
    A = (B + B.T + 2.0 * sp.identity(n)).tocsr()
    xs = 1.0 - 4.0 * rng.random(n)
    return A, -(A @ xs)


def bench_sparse(device, cases=((1 << 20, 24, 8),), solvers_=("SPG",)):
    """Row f-3: whole solves with an operator-form (CSR) Hessian.  Algorithmic bytes per mat-vec =
    12 nnz + 8 (n+1) + 16 n (values + column ids once, row pointers, v in, y out)."""
    import torch
    import problems as pr
    from helpers import make_solver
    from ccqppy_b200 import solution_spaces as ss
    peak, _ = measured_peak()
    out = {}
    ids = {"BBPGD": pr.BBPGD, "SPG": pr.SPG, "MPRGP": pr.MPRGP}
    for n, band, extra in cases:
        A, b = sparse_matrix(n, band, extra)
        At = torch.sparse_csr_tensor(torch.from_numpy(A.indptr.astype(np.int64)), torch.from_numpy(A.indices.astype(np.int64)),
                                     torch.from_numpy(A.data), size=(n, n)).to(device)
        bt = torch.from_numpy(b).to(device)
        op = ss.BoxProjOp(n)
        g = torch.Generator(device=device).manual_seed(3)
        uni = torch.rand(2000, generator=g, dtype=torch.float64, device=device)
        row = {}
        for name in solvers_:
            for _ in range(2):
                s = make_solver(ids[name], 1e-6, 2000)
                s.solve(At, bt, convex_proj_op=op, uniforms=uni)
            gbps = s.solution_hbm_bytes / s.solution_gpu_time / 1e9
            row[name] = dict(mv=int(s.solution_num_matrix_vector_multiplications), gemv=int(s.solution_gemv_count),
                             converged=bool(s.solution_converged), kernel_ms=1e3 * s.solution_gpu_time, GBps=gbps,
                             frac=gbps / peak, us_per_matvec=1e6 * s.solution_gpu_time / s.solution_gemv_count)
        row["nnz"] = int(A.nnz)
        out["n=%d nnz/row=%.1f" % (n, A.nnz / n)] = row
        del At
    return out


def timed_solves(step, steps, sync_all):
    import torch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_s, mvs, launches = 0.0, 0, 0
    sync_all()
    ev0.record()
    for _ in range(steps):
        r = step()
        kern_s += r.solution_gpu_time
        mvs += r.solution_gemv_count
        launches += r.solution_kernel_launches
    ev1.record()
    sync_all()
    return r, ev0.elapsed_time(ev1) * 1e-3, kern_s, mvs, launches


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-batched", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip every CPU leg (parity against the port included)")
    ap.add_argument("--no-sparse", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    capture_stdout()
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from ccqppy_b200 import solvers, solution_spaces as ss
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU path"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    n = N_DENSE
    t_gen = time.time()
    A, b = make_dense_problem(n, device)
    torch.cuda.synchronize()
    log("[bench] generated n=%d problem in %.1f s" % (n, time.time() - t_gen))
    op = ss.BoxProjOp(n)
    uni_host = spg_uniform_stream(MAX_MV)
    uni_dev = torch.from_numpy(uni_host).to(device)
    peak, peak_src = measured_peak()
    rows_local = n // world
    bytes_per_mv = 8.0 * rows_local * n + 8.0 * n + 8.0 * rows_local

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    single_ref = None                    # N > 1: rank 0's single-GPU answers on the same problem, for `parity`
    if world > 1:
        from ccqppy_b200 import dist as cdist
        if rank == 0:
            single_ref = {}
            for name, cls in (("SPG", solvers.CCQPSolverSPG), ("APGD", solvers.CCQPSolverAPGD)):
                s1 = cls(TOL, MAX_MV)
                s1.quiet = True
                s1.solve(A, b, convex_proj_op=op, uniforms=uni_dev)
                single_ref[name] = (s1.solution.cpu().numpy(), int(s1.solution_num_matrix_vector_multiplications),
                                    bool(s1.solution_converged))
        runner = cdist.ShardedSolver(solvers.CCQPSolverSPG(TOL, MAX_MV), A, op, rank, world, device)
        runner.shard = runner.shard.clone()      # keep only this rank's rows resident
        runner.set_matrix(runner.shard)
        del A
        torch.cuda.empty_cache()

        def device_step():
            return runner.solve(b, uniforms=uni_dev)
    else:
        spg = solvers.CCQPSolverSPG(TOL, MAX_MV)
        spg.quiet = True

        def device_step():
            spg.solve(A, b, convex_proj_op=op, uniforms=uni_dev)
            return spg

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        r = device_step()
    sync_all()
    sampler.mark()
    r, dt, kern_s, mvs, launches = timed_solves(device_step, args.steps, sync_all)
    clocks = sampler.stop()
    if world > 1:
        tt = torch.tensor([dt, kern_s], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, kern_s = float(tt[0]), float(tt[1])
    value = mvs / dt
    mv_reported = int(r.solution_num_matrix_vector_multiplications)
    x_spg = r.solution.cpu().numpy() if hasattr(r.solution, "cpu") else np.asarray(r.solution)
    conv_spg = bool(r.solution_converged)
    achieved = (mvs * bytes_per_mv) / kern_s / 1e9
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=None,
                    kernel="dense_kernel<SPG> (one persistent launch per solve)", peak_source=peak_src,
                    algorithmic_bytes_per_launch=(mvs / args.steps) * bytes_per_mv,
                    avg_launch_ms=1e3 * kern_s / args.steps, per_gpu=True, kernel_source_sha256=kernel_source_hash())
    traffic_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(traffic_file) and world == 1:
        try:
            tf = json.load(open(traffic_file))
            if tf.get("kernel_source_sha256") == roofline["kernel_source_sha256"] and tf.get("n") == n:
                roofline["traffic"] = tf.get("dense_spg_bytes_per_launch")
                roofline["traffic_source"] = tf.get("source")
            else:     # an ncu capture of an OLDER build of the kernels (or of another n) says nothing about this one
                roofline["traffic_note"] = ("profiles/dram_traffic.json was captured for kernel sources %s, n=%s; this build is %s"
                                            % (tf.get("kernel_source_sha256"), tf.get("n"), roofline["kernel_source_sha256"]))
        except Exception:
            pass

    # ---- config 3's second solver: APGD on the same problem (one warm-up, then timed solves)
    def apgd_step():
        if world > 1:
            return apgd_runner.solve(b)
        apgd.solve(A, b, convex_proj_op=op)
        return apgd
    if world > 1:
        runner.solver = solvers.CCQPSolverAPGD(TOL, MAX_MV)
        apgd_runner = runner
    else:
        apgd = solvers.CCQPSolverAPGD(TOL, MAX_MV)
        apgd.quiet = True
    apgd_step()
    a_steps = max(2, min(args.steps, 4))
    ra, a_dt, a_kern, a_mvs, a_launches = timed_solves(apgd_step, a_steps, sync_all)
    if world > 1:
        tt = torch.tensor([a_dt, a_kern], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        a_dt, a_kern = float(tt[0]), float(tt[1])
        runner.solver = solvers.CCQPSolverSPG(TOL, MAX_MV)
    x_apgd = ra.solution.cpu().numpy() if hasattr(ra.solution, "cpu") else np.asarray(ra.solution)
    a_gbps = a_mvs * bytes_per_mv / a_kern / 1e9
    apgd_obj = dict(solver="CCQPSolverAPGD", value=a_mvs / a_dt, unit="mat-vecs/s", steps=a_steps, ms_per_solve=1e3 * a_dt / a_steps,
                    mat_vecs_per_solve=a_mvs / a_steps, reported_mv=int(ra.solution_num_matrix_vector_multiplications),
                    converged=bool(ra.solution_converged), GBps_per_gpu=a_gbps, frac=a_gbps / peak, gpu_launches=a_launches)
    launches += a_launches

    line = dict(metric="spg_iterations_per_s_dense_n%d_fp64" % n, value=value, unit="iterations/s", n_gpus=world,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True,
                scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=workload_config(n, world),
                solve=dict(mat_vecs_per_solve=mvs / args.steps, reported_mv=mv_reported, converged=conv_spg),
                clocks=clocks, gpu_launches=launches, roofline=roofline, apgd=apgd_obj,
                effective_GBps_whole_job=value * (8.0 * n * n + 16.0 * n) / 1e9)

    if world == 1:
        # ---- end to end through the public API from pinned host buffers
        t_pin = time.time()
        A_host = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
        A_host.copy_(A)
        b_host = b.cpu().pin_memory()
        uni_pinned = torch.from_numpy(uni_host).pin_memory()
        torch.cuda.synchronize()
        log("[bench] pinned host copy of A in %.1f s" % (time.time() - t_pin))
        del A
        torch.cuda.empty_cache()
        # one solve at a time (latency view): upload, solve, download, strictly in sequence
        e2e_solver = solvers.CCQPSolverSPG(TOL, MAX_MV)
        e2e_solver.quiet = True
        e2e_steps = max(2, min(args.steps, 5))
        e2e_solver.solve(A_host, b_host, convex_proj_op=op, uniforms=uni_pinned)     # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e_mvs = 0
        for _ in range(e2e_steps):
            e2e_solver.solve(A_host, b_host, convex_proj_op=op, uniforms=uni_pinned)
            e_mvs += e2e_solver.solution_gemv_count
        torch.cuda.synchronize()
        seq_dt = time.perf_counter() - t0
        seq_value = e_mvs / seq_dt
        launches += e2e_steps + 1
        # a stream of solves (throughput view): same work per step, but the upload of step k+1 (copy engine)
        # overlaps the solver kernel of step k (SMs): two handles / streams, ccqp_solve_async + ccqp_solve_wait
        from ccqppy_b200.pipeline import SolvePipeline
        del e2e_solver
        from ccqppy_b200 import _capi as _c
        for hd in list(_c._default.values()):      # the synchronous path's handle holds its own 8.6 GB copy of A
            hd.close()
        _c._default.clear()
        torch.cuda.empty_cache()
        pipe = SolvePipeline(solvers.CCQPSolverSPG(TOL, MAX_MV), depth=2, device=local_rank)
        for _ in range(2):
            pipe.submit(A_host, b_host, convex_proj_op=op, uniforms=uni_pinned)
        pipe.results()                              # warm-up: both slots have their buffers
        torch.cuda.synchronize()
        p_steps = max(4, min(args.steps, 12))      # the stream is as long as the resident-A run (fill and drain of the pipeline are inside the timed region)
        t0 = time.perf_counter()
        for _ in range(p_steps):
            pipe.submit(A_host, b_host, convex_proj_op=op, uniforms=uni_pinned)
        res = pipe.results()
        e_dt = time.perf_counter() - t0
        p_mvs = sum(r.solution_gemv_count for r in res)
        assert len(res) == p_steps and all(r.solution_converged for r in res)
        up_bytes, up_mirrored = pipe.slots[0].handle.upload_info()
        # the same stream for a caller that DECLARES the symmetry (solve(..., symmetric=True): no host-side test)
        t0 = time.perf_counter()
        for _ in range(p_steps):
            pipe.submit(A_host, b_host, convex_proj_op=op, uniforms=uni_pinned, symmetric=True)
        res_d = pipe.results()
        d_dt = time.perf_counter() - t0
        d_mvs = sum(r.solution_gemv_count for r in res_d)
        declared = dict(value=d_mvs / d_dt, s_per_solve=d_dt / p_steps, steps=p_steps, h2d_matrix_bytes=pipe.slots[0].handle.upload_info()[0],
                        same_solution=bool(all(np.array_equal(np.asarray(r.solution), np.asarray(res[0].solution)) for r in res_d)),
                        how="submit(..., symmetric=True) -> ccqp_set_matrix_symmetric: the upper block triangle alone is read and uploaded")
        pipe.close()
        launches += 2 * p_steps + 2 + (2 * p_steps + 2 if up_mirrored else p_steps)    # solver kernels (+ mirror kernels)
        if up_mirrored:
            launches += e2e_steps + 1           # the mirror kernels of the one-at-a-time loop above
        line["e2e"] = dict(value=p_mvs / e_dt, unit="iterations/s", h2d_bytes_per_step=up_bytes + 8 * n + 8 * MAX_MV,
                           matrix_upload=dict(bytes=up_bytes, full_bytes=8 * n * n, mirrored_on_device=up_mirrored,
                                              how="ccqp_set_matrix: the upper block triangle of a symmetric A crosses PCIe, host threads verify "
                                                  "the symmetry meanwhile (timed), a kernel mirrors it; any other A is uploaded whole"),
                           d2h_bytes_per_step=8 * n + 72, steps=p_steps, s_per_solve=e_dt / p_steps, mode="pipelined stream of solves",
                           one_at_a_time=dict(value=seq_value, s_per_solve=seq_dt / e2e_steps, steps=e2e_steps),
                           declared_symmetric=declared,
                           note="every solve re-uploads the Hessian from pinned host memory (PCIe bound; see matrix_upload); `value` is a "
                                "THROUGHPUT figure: a stream of solves through ccqppy_b200.pipeline.SolvePipeline (upload of the next "
                                "problem overlaps the current solve); `one_at_a_time` is the plain solve() loop (latency view)")
        line["gpu_launches"] = launches
        # ---- CPU port of the reference on this host's cores: the WHOLE solve, for parity AND as the cpu baseline
        if not args.no_cpu_baseline:
            from threadpoolctl import threadpool_info, threadpool_limits
            from oracle import ccqp_oracle as orc
            import problems as pr
            A_np, b_np = A_host.numpy(), b_host.numpy()
            tab = pr.box_table(n)
            avail = host_threads()
            with threadpool_limits(limits=avail, user_api="blas"):
                cores = max([i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"] or [1])
                orc.solve(orc.SPG, A_np, b_np, blocks=tab.blocks, params=tab.params, tol=TOL, max_mv=4, uniforms=uni_host)
                t0 = time.perf_counter()
                o = orc.solve(orc.SPG, A_np, b_np, blocks=tab.blocks, params=tab.params, tol=TOL, max_mv=MAX_MV, uniforms=uni_host)
                c_dt = time.perf_counter() - t0
                oa = orc.solve(orc.APGD, A_np, b_np, blocks=tab.blocks, params=tab.params, tol=TOL, max_mv=MAX_MV)
            line["cpu_baseline"] = dict(value=o["gemv"] / c_dt, unit="iterations/s", cores=cores, host_threads=avail, kind="port",
                                        seconds=c_dt,
                                        sample="the whole n=%d SPG solve (%d mat-vecs), NumPy/OpenBLAS port of the reference "
                                               "(oracle/, bit-identical to the reference in the build container)" % (n, o["gemv"]))
            line["parity"] = parity_record("oracle/ (CPU port of the reference), whole solve on the same arrays", x_spg, mv_reported,
                                           conv_spg, o["solution"], o["mv"], o["converged"])
            line["apgd"]["parity"] = parity_record("oracle/ (CPU port of the reference), whole solve on the same arrays", x_apgd,
                                                   apgd_obj["reported_mv"], apgd_obj["converged"], oa["solution"], oa["mv"],
                                                   oa["converged"])
            del A_np
        del A_host
        if not args.no_batched:
            try:
                line["batched"] = bench_batched(device, steps=3, warmup=2)
                line["batched"]["workload"] = "%d box-QPs n=%d, A=GG^T/n+I, tol 1e-8, persistent per-CTA kernel" % (BATCH, NB)
                line["gpu_launches"] += 2 * 5 + 8
                if not args.no_cpu_baseline:
                    line["batched"]["cpu_baseline"] = batched_cpu_baseline()
            except Exception as ex:   # never lose the headline line
                line["batched"] = dict(error=repr(ex))
        if not args.no_sparse:
            try:
                line["sparse"] = bench_sparse(device)
                line["sparse"]["workload"] = ("operator-form (CSR) Hessian, banded + random, Box, tol 1e-6, whole SPG solve in the "
                                              "persistent kernel; bytes per mat-vec = 12 nnz + 8 (n+1) + 16 n")
                line["gpu_launches"] += 2
            except Exception as ex:
                line["sparse"] = dict(error=repr(ex))
    else:
        # ---- parity at N GPUs: against rank 0's single-GPU solve of the same problem, and all ranks identical
        gathered = [torch.empty(n, dtype=torch.float64, device=device) for _ in range(world)]
        dist.all_gather(gathered, r.solution.contiguous())
        ident = all(bool(torch.equal(g_, gathered[0])) for g_ in gathered)
        if rank == 0:
            xs_, mv_, cv_ = single_ref["SPG"]
            line["parity"] = parity_record("the single-GPU kernel's solve of the same problem (rank 0)", x_spg, mv_reported, conv_spg,
                                           xs_, mv_, cv_)
            line["parity"]["ranks_identical"] = ident
            line["parity"]["ok"] = line["parity"]["ok"] and ident
            xa_, mva_, cva_ = single_ref["APGD"]
            line["apgd"]["parity"] = parity_record("the single-GPU kernel's solve of the same problem (rank 0)", x_apgd,
                                                   apgd_obj["reported_mv"], apgd_obj["converged"], xa_, mva_, cva_)
        # ---- end to end at N GPUs: every rank re-uploads ITS rows of A from pinned host memory each step
        # (N PCIe links in parallel), b and the uniforms come from pinned host memory, x goes back to the host
        r0, r1 = runner.ranges[rank]
        A_host = torch.empty((r1 - r0, n), dtype=torch.float64, pin_memory=True)
        A_host.copy_(runner.shard)
        b_host = b.cpu().pin_memory()
        uni_pinned = torch.from_numpy(uni_host).pin_memory()
        x_host = torch.empty(n, dtype=torch.float64, pin_memory=True)
        e2e_steps = max(2, min(args.steps, 5))

        def e2e_step():
            runner.set_matrix(A_host)
            r = runner.solve(b_host, uniforms=uni_pinned)
            x_host.copy_(r.solution)
            torch.cuda.synchronize()
            return r
        e2e_step()
        sync_all()
        t0 = time.perf_counter()
        e_mvs = 0
        for _ in range(e2e_steps):
            e_mvs += e2e_step().solution_gemv_count
        sync_all()
        tt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e_dt = float(tt[0])
        line["gpu_launches"] = launches + e2e_steps + 1
        if not args.no_batched:
            del runner, A_host
            torch.cuda.empty_cache()
            from ccqppy_b200.dist import batch_range

            def reduce_max(t, hbm, flops, cnt):
                tt = torch.tensor([t], device=device, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ss_ = torch.tensor([hbm, flops, cnt], device=device, dtype=torch.float64)
                dist.all_reduce(ss_, op=dist.ReduceOp.SUM)
                return float(tt[0]), float(ss_[0]), float(ss_[1]), float(ss_[2]), world
            i0, i1 = batch_range(BATCH, rank, world)
            try:
                line["batched"] = bench_batched(device, steps=3, warmup=2, batch=i1 - i0, seed=1 + rank, reduce_max=reduce_max)
                line["batched"]["workload"] = ("%d box-QPs n=%d split contiguously over %d GPUs (no communication), rates for the "
                                               "whole job, fractions against %d x the per-GPU peaks" % (BATCH, NB, world, world))
                line["gpu_launches"] += 2 * 5 + 8
            except Exception as ex:
                line["batched"] = dict(error=repr(ex))
        line["e2e"] = dict(value=e_mvs / e_dt, unit="iterations/s",
                           h2d_bytes_per_step=8 * n * n + world * (8 * n + 8 * MAX_MV), d2h_bytes_per_step=world * (8 * n + 72),
                           steps=e2e_steps, s_per_solve=e_dt / e2e_steps, mode="one solve at a time",
                           note="every solve re-uploads the Hessian: each rank copies its %d rows (%.2f GB) from pinned host "
                                "memory over its own PCIe link" % (r1 - r0, 8e-9 * (r1 - r0) * n))
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
