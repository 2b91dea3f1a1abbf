// CPU emulation of the register-resident three-stage FFT of fastpm_b200/csrc/fft_reg.cuh (struct Fft3), the core of the TMA tile
// pass (fft_tma.cu) and of the row pass (fft_zrow.cu).  Every CUDA thread of one CTA becomes an OS thread, __syncthreads() a
// pthread barrier, shared memory a plain array; the thread <-> element mapping, the swizzled exchange buffer and the twiddle
// indexing are exactly the device code.  Checked against a naive double-precision DFT.  Prints "OK" lines; exit code 0 on success.
#include <pthread.h>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

static pthread_barrier_t g_barrier;
#define __syncthreads() pthread_barrier_wait(&g_barrier)
#define __ldg(p) (*(p))
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t) p; }      // only used by the PTX wrappers (never called here)
#include "../../fastpm_b200/csrc/fft_reg.cuh"

typedef std::complex<double> cd;

template <int R1, int R2, int R3, int K, bool TWS>
static double run_config(const char *name)
{
    using C = TmaCfg<R1, R2, R3>;
    constexpr int N = C::N, E = C::E, T = C::T, M1 = C::M1;
    const int nthreads = T * K;
    std::vector<float2> in((size_t) N * K), out((size_t) N * K), tw(N);
    std::vector<float> B((size_t) N * K);
    for (int t = 0; t < N; t++) tw[t] = make_float2((float) cos(-2.0 * M_PI * t / N), (float) sin(-2.0 * M_PI * t / N));
    srand(N * 31 + K);
    for (auto &v : in) v = make_float2(rand() / (float) RAND_MAX - 0.5f, rand() / (float) RAND_MAX - 0.5f);
    pthread_barrier_init(&g_barrier, NULL, nthreads);
    std::vector<std::thread> pool;
    for (int tid = 0; tid < nthreads; tid++) {
        pool.emplace_back([&, tid]() {
            const int c = tid % K, t = tid / K;
            Fft3<R1, R2, R3, K, TWS> fx(B.data(), t, c, tw.data());
            float2 v[E];
            for (int k = 0; k < E; k++) v[k] = in[(size_t) (t + k * M1) * K + c];           // element t + k*M1 of column c
            fx.run(v, []() {});
            for (int i = 0; i < E / R3; i++) {
                const int b = t + i * T, q1 = b / R2, q2 = b - q1 * R2;
                for (int q3 = 0; q3 < R3; q3++) out[(size_t) (q1 + R1 * q2 + R1 * R2 * q3) * K + c] = v[i * R3 + q3];
            }
        });
    }
    for (auto &th : pool) th.join();
    pthread_barrier_destroy(&g_barrier);
    double maxerr = 0, norm = 0;
    for (int c = 0; c < K; c += (K > 1 ? K - 1 : 1)) {
        for (int k = 0; k < N; k += (N > 256 ? 7 : 1)) {
            cd acc = 0;
            for (int e = 0; e < N; e++)
                acc += cd(in[(size_t) e * K + c].x, in[(size_t) e * K + c].y) * std::polar(1.0, -2.0 * M_PI * (double) ((long long) e * k % N) / N);
            const cd got(out[(size_t) k * K + c].x, out[(size_t) k * K + c].y);
            maxerr = std::max(maxerr, std::abs(got - acc));
            norm = std::max(norm, std::abs(acc));
        }
    }
    const double rel = maxerr / norm;
    printf("Fft3 %-18s N=%4d K=%2d threads=%4d relerr=%.2e %s\n", name, N, K, nthreads, rel, rel < 2e-6 ? "OK" : "FAIL");
    return rel;
}

int main()
{
    double worst = 0;
    worst = std::max(worst, run_config<8, 8, 8, 16, true>("tile 512"));
    worst = std::max(worst, run_config<16, 16, 4, 16, true>("tile 1024"));
    worst = std::max(worst, run_config<16, 16, 8, 8, true>("tile 2048"));
    worst = std::max(worst, run_config<16, 16, 16, 4, true>("tile 4096"));
    worst = std::max(worst, run_config<8, 8, 4, 8, false>("row 512 (H=256)"));
    worst = std::max(worst, run_config<8, 8, 8, 8, false>("row 1024 (H=512)"));
    worst = std::max(worst, run_config<16, 16, 4, 8, false>("row 2048 (H=1024)"));
    worst = std::max(worst, run_config<16, 16, 8, 8, false>("row 4096 (H=2048)"));
    return worst < 2e-6 ? 0 : 1;
}
