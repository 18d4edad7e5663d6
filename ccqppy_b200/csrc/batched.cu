// Translation unit of the batched solver kernels (batched.cuh); see internal.h.
#include "batched.cuh"
#include "internal.h"

namespace ccqp {

int batched_solve_entry(cudaStream_t stream, int sm_count, int solver, const ccqp_params& prm, long long batch, long long n,
                        const double* A, const double* b, const double* x0, const double* lb, const double* ub,
                        const double* uniforms, long long n_uniforms, double* x_out, int memtype, ccqp_result* results,
                        ccqp_result* summary, cudaEvent_t ev0, cudaEvent_t ev1, int* launches, std::string& err,
                        const std::function<void*(size_t)>& alloc) {
    return batched_solve(stream, sm_count, nullptr, 0, solver, prm, batch, n, A, b, x0, lb, ub, uniforms, n_uniforms, x_out,
                         memtype, results, summary, ev0, ev1, launches, err, alloc);
}

}  // namespace ccqp
