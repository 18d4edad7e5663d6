"""BASELINE.json's FULL sizes on the GPU.

Two kinds of checks:
  * against the oracle (the CPU port of the reference) on the SAME arrays -- the problem is generated on the GPU
    and copied to the host; the port needs about 5 s for the 56-mat-vec SPG solve at n = 32768 on the box's host
    cores and a few seconds for MPRGP at n = 16384 (north_star's tolerance: solution within 1e-9 relative, the
    same converged flag, mat-vec count within 2 %);
  * size-independent properties: feasibility, the reference's own fixed-point residual recomputed independently
    with torch, agreement between different solvers on the same strictly convex problem, run-to-run
    determinism, and the mat-vec accounting (the only checks possible for the 65536-problem batch, which the
    one-problem-at-a-time port would need minutes for; a sample of it is compared problem by problem)."""
import numpy as np
import pytest

import problems as pr
from helpers import op_from_table, make_solver

pytestmark = pytest.mark.gpu


def gpu_problem(n, seed=0):
    """Same generator as bench.py: A = G G^T / n + I on the device, b = -A x*."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    G = torch.randn((n, n), generator=g, device="cuda", dtype=torch.float64)
    A = G @ G.t()
    del G
    A.div_(n)
    A.diagonal().add_(1.0)
    xs = 1.0 - 4.0 * torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
    return A, -(A @ xs)


def residual(A, b, x, project):
    """RES of solvers.py:137-139 recomputed with torch: |x - P(x - 1e-6 g)| / (3 n 1e-6)."""
    g = A @ x + b
    return float((x - project(x - 1e-6 * g)).norm()) / (3 * x.numel() * 1e-6)


def assert_matches_oracle(s, o, what):
    """north_star: solution within 1e-9 relative, same converged flag, mat-vec count within 2 %."""
    x = s.solution.cpu().numpy() if hasattr(s.solution, "cpu") else np.asarray(s.solution)
    mv = int(s.solution_num_matrix_vector_multiplications)
    assert bool(s.solution_converged) == bool(o["converged"]), what
    assert abs(mv - o["mv"]) <= max(1, round(0.02 * o["mv"])), (what, mv, o["mv"])
    rel = np.linalg.norm(x - o["solution"]) / np.linalg.norm(o["solution"])
    assert rel <= 1e-9, (what, rel)
    return mv, rel


def test_config3_spg_and_apgd_n32768_match_oracle():
    """BASELINE config 3 at full size against the port: solvers.py:878-975 (SPG) and :220-343 (APGD)."""
    import torch
    from oracle import ccqp_oracle as orc
    from ccqppy_b200 import solution_spaces as ss
    n, tol = 32768, 1e-5
    A, b = gpu_problem(n)
    op = ss.BoxProjOp(n)
    uni = pr.spg_uniforms(0, 2000)
    spg = make_solver(pr.SPG, tol, 2000)
    spg.solve(A, b, convex_proj_op=op, uniforms=torch.from_numpy(uni).cuda())
    apgd = make_solver(pr.APGD, tol, 2000)
    apgd.solve(A, b, convex_proj_op=op)
    A_np, b_np = A.cpu().numpy(), b.cpu().numpy()
    tab = pr.box_table(n)
    o = orc.solve(orc.SPG, A_np, b_np, blocks=tab.blocks, params=tab.params, tol=tol, max_mv=2000, uniforms=uni)
    mv, rel = assert_matches_oracle(spg, o, "SPG n=32768")
    assert mv == o["mv"]                       # on the well-conditioned generator the counts are in fact identical
    o = orc.solve(orc.APGD, A_np, b_np, blocks=tab.blocks, params=tab.params, tol=tol, max_mv=2000)
    assert_matches_oracle(apgd, o, "APGD n=32768")


@pytest.mark.parametrize("table", ["sphere3", "mixed"])
def test_config5_mprgp_n16384_matches_oracle(table):
    """BASELINE config 5 at full size against the port: solvers.py:1026-1200 with 5461 Sphere(3) friction discs
    (solution_spaces.py:369-435) / the mixed disjoint table (every working operator kind, :495-560)."""
    from oracle import ccqp_oracle as orc
    n, tol = 16384, 1e-5
    A, b = gpu_problem(n, seed=3)
    tab = {"sphere3": pr.sphere3_table, "mixed": pr.mixed_table}[table](n)
    s = make_solver(pr.MPRGP, tol, 3000)
    s.solve(A, b, convex_proj_op=op_from_table(tab))
    o = orc.solve(orc.MPRGP, A.cpu().numpy(), b.cpu().numpy(), blocks=tab.blocks, params=tab.params, tol=tol, max_mv=3000)
    assert_matches_oracle(s, o, "MPRGP n=16384 " + table)


def test_config3_dense_spg_n32768_properties():
    import torch
    from ccqppy_b200 import solution_spaces as ss
    n, tol = 32768, 1e-5
    A, b = gpu_problem(n)
    op = ss.BoxProjOp(n)
    box = lambda v: v.clamp(-1.0, 1.0)
    uni = torch.from_numpy(pr.spg_uniforms(0, 2000)).cuda()
    outs = []
    for _ in range(2):
        s = make_solver(pr.SPG, tol, 2000)
        s.solve(A, b, convex_proj_op=op, uniforms=uni)
        outs.append(s)
    s = outs[0]
    x = s.solution
    assert s.solution_converged and s.solution_kernel_launches == 1
    assert s.solution_gemv_count == s.solution_num_matrix_vector_multiplications   # SPG counts every product
    assert bool((x.abs() <= 1.0).all())
    # SPG stops on |P(x - alpha g) - x| <= tol; the scaled fixed-point residual is then tiny as well
    assert residual(A, b, x, box) < 1e-4
    assert torch.equal(x, outs[1].solution) and s.solution_residual == outs[1].solution_residual   # deterministic
    # a different algorithm finds the same minimiser of the strictly convex problem
    bb = make_solver(pr.BBPGD, 1e-7, 2000)
    bb.solve(A, b, convex_proj_op=op)
    assert bb.solution_converged and residual(A, b, bb.solution, box) < 1e-7 * 1.01
    assert float((bb.solution - x).norm() / x.norm()) < 1e-5
    # roofline accounting: bytes = executed mat-vecs x (8 n^2 + 16 n)
    assert s.solution_hbm_bytes == s.solution_gemv_count * (8.0 * n * n + 16.0 * n)


def test_config5_mprgp_n16384_friction_discs_properties():
    """Contact-style problem: 5461 Sphere(3) blocks (+1 free unknown), MPRGP, n = 16384."""
    import torch
    n, tol = 16384, 1e-5
    A, b = gpu_problem(n, seed=3)
    tab = pr.sphere3_table(n)
    op = op_from_table(tab)

    def discs(v):
        out = v.clone()
        blk = out[:3 * (n // 3)].view(-1, 3)
        r = blk.norm(dim=1, keepdim=True)
        blk.copy_(torch.where(r > 1.0, blk / r, blk))
        return out
    s = make_solver(pr.MPRGP, tol, 3000)
    s.solve(A, b, convex_proj_op=op)
    x = s.solution
    assert s.solution_converged
    assert float(x[:3 * (n // 3)].view(-1, 3).norm(dim=1).max()) <= 1.0 + 1e-12
    assert residual(A, b, x, discs) < tol * 1.01
    assert abs(residual(A, b, x, discs) - s.solution_residual) <= 1e-3 * s.solution_residual + 1e-12
    # the lazily evaluated alpha_bb products are never needed on this path: 3 executed per iteration
    assert s.solution_gemv_count <= s.solution_num_matrix_vector_multiplications
    # the reference's residual is scaled by 1/(3n), so tol = 1e-5 is a loose test at this size; two
    # different algorithms at a tight tolerance must meet at the same minimiser
    tight = make_solver(pr.MPRGP, 1e-8, 5000)
    tight.solve(A, b, convex_proj_op=op)
    other = make_solver(pr.BBPGD, 1e-9, 5000)
    other.solve(A, b, convex_proj_op=op)
    assert tight.solution_converged and other.solution_converged
    assert float((other.solution - tight.solution).norm() / x.norm()) < 1e-4
    # the device projection hook agrees with the torch restatement on a random vector
    v = 2 * torch.randn(n, device="cuda", dtype=torch.float64)
    np.testing.assert_allclose(np.asarray(op(v.cpu().numpy())), discs(v).cpu().numpy(), rtol=1e-15, atol=1e-300)


def test_config4_batched_65536_properties():
    import torch
    batch, n, tol = 65536, 64, 1e-8
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.empty((batch, n, n), device="cuda", dtype=torch.float64)
    for s0 in range(0, batch, 16384):
        G = torch.randn((16384, n, n), generator=g, device="cuda", dtype=torch.float64)
        A[s0:s0 + 16384] = G @ G.transpose(1, 2) / n + torch.eye(n, device="cuda", dtype=torch.float64)
    xs = 1 - 4 * torch.rand((batch, n), generator=g, device="cuda", dtype=torch.float64)
    b = -(A @ xs.unsqueeze(-1)).squeeze(-1)
    lb, ub = -torch.ones_like(b), torch.ones_like(b)
    sol = {}
    for solver in (pr.BBPGD, pr.SPG):
        s = make_solver(solver, tol, 5000)
        s.solve_batched(A, b, lb, ub, n_uniforms=256)
        x = s.solution
        assert s.solution_converged.all() and s.solution_kernel_launches == 1
        assert bool(((x >= lb) & (x <= ub)).all())
        grad = (A @ x.unsqueeze(-1)).squeeze(-1) + b
        if solver == pr.BBPGD:
            res = (x - torch.clamp(x - 1e-6 * grad, lb, ub)).norm(dim=1) / (3 * n * 1e-6)
            assert float(res.max()) < tol * 1.001
            np.testing.assert_allclose(res.cpu().numpy(), s.solution_residual, rtol=1e-3, atol=1e-13)
        sol[solver] = x
        if solver == pr.BBPGD:          # a sample of the batch against the port, problem by problem
            from oracle import ccqp_oracle as orc
            tab = pr.box_table(n)
            mvs = s.solution_num_matrix_vector_multiplications
            for i in range(0, batch, 4099):
                o = orc.solve(orc.BBPGD, A[i].cpu().numpy(), b[i].cpu().numpy(), blocks=tab.blocks, params=tab.params,
                              tol=tol, max_mv=5000)
                assert abs(int(mvs[i]) - o["mv"]) <= max(1, round(0.02 * o["mv"])) and o["converged"]
                if int(mvs[i]) == o["mv"]:
                    assert np.linalg.norm(x[i].cpu().numpy() - o["solution"]) <= 1e-9 * np.linalg.norm(o["solution"])
    rel = (sol[pr.BBPGD] - sol[pr.SPG]).norm(dim=1) / sol[pr.BBPGD].norm(dim=1)
    assert float(rel.max()) < 1e-6
