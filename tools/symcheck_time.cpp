// Speed of the host-side symmetry test (csrc/symcheck.h) on this machine's cores:  symcheck_time <n> <threads, 0 = all>
#include "symcheck.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv) {
    long long n = atoll(argv[1]); int th = atoi(argv[2]);
    if (th <= 0) th = ccqp::host_threads_available();
    double* A = (double*)malloc(n * n * 8);
    for (long long i = 0; i < n; ++i) for (long long j = 0; j <= i; ++j) { double v = (double)((i * 131 + j * 7) % 1000) * 1e-3; A[i*n+j] = v; A[j*n+i] = v; }
    for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        bool r = ccqp::host_lower_blocks_mirror_upper(A, n, n, 1024, th);
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("n %lld threads %d sym %d %.1f ms %.1f GB/s\n", n, th, (int)r, dt * 1e3, n * n * 8 / dt / 1e9);
    }
}
