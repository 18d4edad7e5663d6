// Host -> device upload of a dense Hessian that turns out to be symmetric: half of the PCIe traffic.
//
// solve() from host memory is PCIe-bound for large problems (n = 32768: 8.59 GB, ~156 ms at 55 GB/s against a 70 ms solve;
// DESIGN.md section 5).  The reference takes ANY square A (solvers.py:94, A.dot(v) at :133), so nothing may be assumed -- but a
// QP Hessian is symmetric in practice, and whether this one is can be decided on the host while the copy engine is busy:
//   1. the upper block triangle (row blocks of kUpBlock rows, each from its first column on) is enqueued on the handle's stream
//      as 2-D copies straight out of the caller's buffer -- 0.5 n^2 + 0.5 n kUpBlock entries;
//   2. meanwhile host threads compare A[i][j] with A[j][i] for every pair the mirror step would fill in, tile by tile (two
//      32 KB tiles per step, both cache resident), stopping at the first difference (NaNs count as different);
//   3. symmetric: one kernel mirrors the uploaded part into the blocks below the diagonal -- the device copy is then
//      BIT-IDENTICAL to a full upload.  (The kernel is launched by the first user of the matrix, AFTER that user's own small
//      host -> device copies: a kernel between two groups of copies of one stream kept the solver kernel behind the copies
//      another stream enqueued later -- a stream of solves lost the overlap of upload k+1 with solve k; tools/e2e_timeline.py); not symmetric: the missing blocks are uploaded after all (the full copy, in two parts).
// Either way the solver sees exactly the matrix the caller passed.  CCQP_SYM_UPLOAD=0 turns the scheme off.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "internal.h"
#include "symcheck.h"

namespace ccqp {

namespace {

constexpr int kUpBlock = 1024;      // rows per uploaded block (multiple of the mirror tile and of 4: 32-byte aligned row pieces)
constexpr int kMirTile = 32;

// dst[i][j] = dst[j][i] for every j < kUpBlock * (i / kUpBlock): the blocks below the block diagonal from the ones above
__global__ void __launch_bounds__(256) mirror_lower_kernel(double* __restrict__ A, long long n, long long ld) {
    __shared__ double tile[kMirTile][kMirTile + 1];
    const long long i0 = (long long)blockIdx.y * kMirTile, j0 = (long long)blockIdx.x * kMirTile;   // destination tile (rows i0.., columns j0..)
    if (j0 >= (i0 / kUpBlock) * kUpBlock) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // source tile: rows j0.., columns i0.. (coalesced along the columns)
#pragma unroll
    for (int r = ty; r < kMirTile; r += 8) {
        const long long sr = j0 + r, sc = i0 + tx;
        tile[r][tx] = (sr < n && sc < n) ? A[sr * ld + sc] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < kMirTile; r += 8) {
        const long long dr = i0 + r, dc = j0 + tx;
        if (dr < n && dc < n) A[dr * ld + dc] = tile[tx][r];
    }
}

}  // namespace

// Uploads the n x n host matrix A (leading dimension lda) to dst (leading dimension ldd, a multiple of 4) on `stream`.
// declared_symmetric: the caller vouches for the symmetry (ccqp_set_matrix_symmetric): no test, the blocks below the block
// diagonal of A are never read.  *used_mirror reports which way it went, *bytes what crossed PCIe.  Returns the first CUDA error.
int upload_block_rows() { return kUpBlock; }

cudaError_t mirror_lower(cudaStream_t stream, double* dst, long long n, long long ldd) {
    const unsigned tiles = (unsigned)((n + kMirTile - 1) / kMirTile);
    mirror_lower_kernel<<<dim3(tiles, tiles), 256, 0, stream>>>(dst, n, ldd);
    return cudaGetLastError();
}

bool host_matrix_mirrors(const double* A, long long n, long long lda, int threads) {
    return host_lower_blocks_mirror_upper(A, n, lda, kUpBlock, threads > 0 ? threads : host_threads_available());
}

cudaError_t upload_square_matrix(cudaStream_t stream, double* dst, long long ldd, const double* A, long long n, long long lda,
                                 bool declared_symmetric, bool* used_mirror, long long* bytes) {
    *used_mirror = false;
    *bytes = n * n * 8;
    const char* env = getenv("CCQP_SYM_UPLOAD");
    // The test reads all of A from host memory while the copy engine reads half of it: it pays off where the host has the
    // cores (and with them the memory bandwidth) to do that faster than PCIe moves the other half -- measured on a 16-thread
    // host: 70 ms alone / ~95 ms next to the copies against 156 ms for the full upload at n = 32768.
    const int threads = host_threads_available();
    // a declared symmetry holds at every size (the blocks below the block diagonal are never read); the test only pays off for
    // matrices of at least two blocks by two
    const bool enabled = declared_symmetric ? n > kUpBlock : (!(env && atoi(env) == 0) && threads >= 12 && n >= 2 * kUpBlock);
    if (!enabled)
        return cudaMemcpy2DAsync(dst, (size_t)ldd * 8, A, (size_t)lda * 8, (size_t)n * 8, (size_t)n, cudaMemcpyHostToDevice, stream);
    cudaError_t e;
    for (long long i0 = 0; i0 < n; i0 += kUpBlock) {         // 1. upper block triangle
        const long long rows = std::min<long long>(kUpBlock, n - i0);
        e = cudaMemcpy2DAsync(dst + i0 * ldd + i0, (size_t)ldd * 8, A + i0 * lda + i0, (size_t)lda * 8, (size_t)(n - i0) * 8,
                              (size_t)rows, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
    }
    if (declared_symmetric || host_lower_blocks_mirror_upper(A, n, lda, kUpBlock, threads)) {  // 2. (runs while the copies are in flight)
        *used_mirror = true;                                                        // 3a. mirror_lower(), launched by the caller
        *bytes = 0;
        for (long long i0 = 0; i0 < n; i0 += kUpBlock) *bytes += std::min<long long>(kUpBlock, n - i0) * (n - i0) * 8;
        return cudaGetLastError();
    }
    for (long long i0 = kUpBlock; i0 < n; i0 += kUpBlock) {  // 3b. the blocks below the diagonal after all
        const long long rows = std::min<long long>(kUpBlock, n - i0);
        e = cudaMemcpy2DAsync(dst + i0 * ldd, (size_t)ldd * 8, A + i0 * lda, (size_t)lda * 8, (size_t)i0 * 8, (size_t)rows,
                              cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace ccqp
