#!/usr/bin/env python
"""Headline benchmark of the CCQP projected-gradient hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference arm)

Workload (BASELINE.json config 3): dense SPG-QP, n = 32768, fp64,
A = G G^T / n + I (seed 0), b = -A x*, x* = 1 - 4 U, Box[-1,1]^n, tol 1e-5, max 2000 mat-vecs,
SPG uniforms = RandomState(0).  One "step" = one whole solve (about 60 mat-vecs, each streaming
the 8.59 GB Hessian once).  metric = SPG iterations (= mat-vecs executed) per second.

  value : device-resident A (already in HBM), timed with CUDA events around the K solves
  e2e   : the same solves through the public API from pinned HOST buffers: the timed region
          includes the host->device copy of A, b, uniforms and the device->host copy of x
  roofline : algorithmic bytes of the solver kernel (mat-vecs x (8 n^2 + 16 n)) / its duration,
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline : the NumPy/OpenBLAS port of the reference (oracle/) on this host's cores, on a
                 bounded number of mat-vecs of the same problem
  batched : config 4 (65536 box-QPs of n = 64, BBPGD and SPG) in the persistent per-CTA kernel

With N > 1 (torchrun, one rank per GPU) A is row-sharded and the SAME problem is solved by all
ranks together (strong scaling); the exchange of the mat-vec input and of the scalar partial sums
happens inside the solver kernel through NVLink peer memory.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_DENSE = int(os.environ.get("CCQP_BENCH_N", 32768))
TOL, MAX_MV, SEED = 1e-5, 2000, 0
BATCH, NB = int(os.environ.get("CCQP_BENCH_BATCH", 65536)), 64
REF_SAMPLE_MV = 8          # mat-vecs per step of the reference arm / cpu baseline sample
CPU_BASELINE_MV = 40


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


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

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
            self.rows.append([c.strip() for c in line.split(",")])

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
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


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
    A.diagonal().add_(1.0)
    xs = 1.0 - 4.0 * torch.rand(n, generator=g, device=device, dtype=torch.float64)
    b = -(A @ xs)
    return A, b


def spg_uniform_stream(count):
    return np.random.RandomState(SEED).random_sample(count)


def run_reference_arm(args, rank):
    """The reference's own algorithm on the host cores: the NumPy/OpenBLAS port in oracle/ (the
    Python reference cannot travel to the GPU box; the port is bit-identical to it in the build
    container, see oracle/gen_golden.py).  Each step = REF_SAMPLE_MV mat-vecs of the workload."""
    if rank != 0:
        return
    import torch
    from threadpoolctl import threadpool_info
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
    cores = max([i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"] or [1])

    def step():
        o = orc.solve(orc.SPG, A, b, blocks=tab.blocks, params=tab.params, tol=TOL, max_mv=REF_SAMPLE_MV, uniforms=uni)
        return o["gemv"]
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    mvs = sum(step() for _ in range(args.steps))
    dt = time.perf_counter() - t0
    value = mvs / dt
    line = dict(metric="spg_iterations_per_s_dense_n%d_fp64" % n, value=value, unit="iterations/s", impl="reference",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * dt / args.steps,
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload="dense SPG-QP n=%d fp64, A=GG^T/n+I seed 0, Box[-1,1], tol 1e-5" % n,
                            sample="%d mat-vecs per step (solver capped with max_mv)" % REF_SAMPLE_MV),
                cpu_baseline=dict(value=value, unit="iterations/s", cores=cores, kind="port",
                                  sample="%d steps x %d mat-vecs of the n=%d SPG solve, NumPy/OpenBLAS port of the "
                                         "reference" % (args.steps, REF_SAMPLE_MV, n)),
                e2e=dict(value=value, unit="iterations/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                effective_GBps=value * (8.0 * n * n + 16.0 * n) / 1e9)
    emit(line)


# ---------------------------------------------------------------------------------------------
def bench_batched(device, steps, warmup, batch=None, seed=1, reduce_max=None):
    """Config 4.  `batch` problems on this GPU; with `reduce_max` (multi-GPU) the slowest rank's time counts and
    the rates are for the whole job of BATCH problems."""
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
        xs = 1 - 4 * torch.rand((e - s, NB), generator=g, device=device, dtype=torch.float64)
        b[s:e] = -(A[s:e] @ xs.unsqueeze(-1)).squeeze(-1)
    lb, ub = -torch.ones_like(b), torch.ones_like(b)
    K = 256
    uni = torch.rand((local, K), generator=g, device=device, dtype=torch.float64)
    peak, _ = measured_peak()
    for name, cls in (("BBPGD", solvers.CCQPSolverBBPGD), ("SPG", solvers.CCQPSolverSPG)):
        s = cls(1e-8, 5000)
        s.quiet = True
        times = []
        for it in range(warmup + steps):
            s.solve_batched(A, b, lb, ub, uniforms=uni)
            if it >= warmup:
                times.append(s.solution_gpu_time)
        t = float(np.mean(times))
        hbm = s.solution_hbm_bytes
        flops = 2.0 * NB * NB * s.solution_gemv_count
        total = local
        if reduce_max is not None:      # slowest rank's time; bytes, flops and problems of all ranks
            t, hbm, flops, total = reduce_max(t, hbm, flops, local)
        out[name] = dict(qps=total / t, ms=1e3 * t, mean_mv=float(np.mean(s.solution_num_matrix_vector_multiplications)),
                         converged=bool(np.all(s.solution_converged)), hbm_GBps=hbm / t / 1e9, hbm_frac=hbm / t / 1e9 / peak,
                         fp64_TFLOPs=flops / t / 1e12)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-batched", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    if world > 1:
        from ccqppy_b200 import dist as cdist
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

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        r = device_step()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_s, mvs, launches = 0.0, 0, 0
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        r = device_step()
        kern_s += r.solution_gpu_time
        mvs += r.solution_gemv_count
        launches += r.solution_kernel_launches
    ev1.record()
    sync_all()
    clocks = sampler.stop()
    dt = ev0.elapsed_time(ev1) * 1e-3
    if world > 1:
        tt = torch.tensor([dt, kern_s], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, kern_s = float(tt[0]), float(tt[1])
    value = mvs / dt
    mv_reported = int(r.solution_num_matrix_vector_multiplications)
    peak, peak_src = measured_peak()
    rows_local = n // world
    bytes_per_mv = 8.0 * rows_local * n + 8.0 * n + 8.0 * rows_local
    achieved = (mvs * bytes_per_mv) / kern_s / 1e9
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=None,
                    kernel="dense_kernel<SPG> (one persistent launch per solve)", peak_source=peak_src,
                    algorithmic_bytes_per_launch=(mvs / args.steps) * bytes_per_mv,
                    avg_launch_ms=1e3 * kern_s / args.steps, per_gpu=True)
    traffic_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(traffic_file) and world == 1:
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get("dense_spg_bytes_per_launch")
        except Exception:
            pass

    line = dict(metric="spg_iterations_per_s_dense_n%d_fp64" % n, value=value, unit="iterations/s", n_gpus=world,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True,
                scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload="dense SPG-QP n=%d fp64, A=GG^T/n+I seed 0, Box[-1,1], tol 1e-5, max_mv 2000; "
                                     "A row-sharded over %d GPU(s)" % (n, world),
                            l2="inputs (%.1f GB per GPU) far exceed the 126 MB L2; no flush needed" % (8e-9 * rows_local * n),
                            mat_vecs_per_solve=mvs / args.steps, reported_mv=mv_reported,
                            converged=bool(r.solution_converged)),
                clocks=clocks, gpu_launches=launches, roofline=roofline,
                effective_GBps_whole_job=value * (8.0 * n * n + 16.0 * n) / 1e9)

    # ---- end to end through the public API from pinned host buffers (rank 0 view; N=1 only copies all of A)
    if world == 1:
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
        p_steps = max(4, e2e_steps)
        t0 = time.perf_counter()
        for _ in range(p_steps):
            pipe.submit(A_host, b_host, convex_proj_op=op, uniforms=uni_pinned)
        res = pipe.results()
        e_dt = time.perf_counter() - t0
        p_mvs = sum(r.solution_gemv_count for r in res)
        assert len(res) == p_steps and all(r.solution_converged for r in res)
        pipe.close()
        line["e2e"] = dict(value=p_mvs / e_dt, unit="iterations/s", h2d_bytes_per_step=8 * n * n + 8 * n + 8 * MAX_MV,
                           d2h_bytes_per_step=8 * n + 72, steps=p_steps, s_per_solve=e_dt / p_steps,
                           one_at_a_time=dict(value=seq_value, s_per_solve=seq_dt / e2e_steps, steps=e2e_steps),
                           note="every solve re-uploads the 8.59 GB Hessian from pinned host memory (PCIe bound); `value` is a "
                                "stream of solves through ccqppy_b200.pipeline.SolvePipeline (upload of the next problem overlaps "
                                "the current solve), `one_at_a_time` the plain solve() loop")
        line["gpu_launches"] = launches
        # ---- CPU baseline: the port of the reference on this host's cores, bounded sample
        if not args.no_cpu_baseline:
            from threadpoolctl import threadpool_info
            from oracle import ccqp_oracle as orc
            import problems as pr
            A_np, b_np = A_host.numpy(), b_host.numpy()
            tab = pr.box_table(n)
            cores = max([i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"] or [1])
            orc.solve(orc.SPG, A_np, b_np, blocks=tab.blocks, params=tab.params, tol=TOL, max_mv=4, uniforms=uni_host)
            t0 = time.perf_counter()
            o = orc.solve(orc.SPG, A_np, b_np, blocks=tab.blocks, params=tab.params, tol=TOL, max_mv=CPU_BASELINE_MV,
                          uniforms=uni_host)
            c_dt = time.perf_counter() - t0
            line["cpu_baseline"] = dict(value=o["gemv"] / c_dt, unit="iterations/s", cores=cores, kind="port",
                                        sample="first %d mat-vecs of the same n=%d SPG solve, NumPy/OpenBLAS port of the "
                                               "reference (os.cpu_count()=%d)" % (o["gemv"], n, os.cpu_count()))
        del A_host
        if not args.no_batched:
            try:
                line["batched"] = bench_batched(device, steps=3, warmup=2)
                line["batched"]["workload"] = "%d box-QPs n=%d, A=GG^T/n+I, tol 1e-8, persistent per-CTA kernel" % (BATCH, NB)
                if not args.no_cpu_baseline:
                    line["batched"]["cpu_baseline"] = batched_cpu_baseline()
            except Exception as ex:   # never lose the headline line
                line["batched"] = dict(error=repr(ex))
    else:
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
        if not args.no_batched:
            del runner, A_host
            torch.cuda.empty_cache()
            from ccqppy_b200.dist import batch_range

            def reduce_max(t, hbm, flops, cnt):
                tt = torch.tensor([t], device=device, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ss = torch.tensor([hbm, flops, cnt], device=device, dtype=torch.float64)
                dist.all_reduce(ss, op=dist.ReduceOp.SUM)
                return float(tt[0]), float(ss[0]), float(ss[1]), float(ss[2])
            i0, i1 = batch_range(BATCH, rank, world)
            try:
                line["batched"] = bench_batched(device, steps=3, warmup=2, batch=i1 - i0, seed=1 + rank, reduce_max=reduce_max)
                line["batched"]["workload"] = ("%d box-QPs n=%d split contiguously over %d GPUs (no communication), rates for the "
                                               "whole job, hbm_frac per job against %d x the per-GPU peak" % (BATCH, NB, world, world))
                for v in line["batched"].values():
                    if isinstance(v, dict):
                        v["hbm_frac"] = v["hbm_frac"] / world
            except Exception as ex:
                line["batched"] = dict(error=repr(ex))
        line["e2e"] = dict(value=e_mvs / e_dt, unit="iterations/s",
                           h2d_bytes_per_step=8 * n * n + world * (8 * n + 8 * MAX_MV), d2h_bytes_per_step=world * (8 * n + 72),
                           steps=e2e_steps, s_per_solve=e_dt / e2e_steps,
                           note="every solve re-uploads the Hessian: each rank copies its %d rows (%.2f GB) from pinned host "
                                "memory over its own PCIe link" % (r1 - r0, 8e-9 * (r1 - r0) * n))
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
