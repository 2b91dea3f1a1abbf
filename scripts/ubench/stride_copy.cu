// Micro-benchmark: how much HBM bandwidth does a tile-shaped strided copy get on B200, as a function of the row stride
// (TLB reach), the contiguous segment size and a blocked layout?  Informs the FFT pass / k-space layout design.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stride_copy stride_copy.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

struct Pat {            // address (in float4 units of 16 B) of row r of tile (o, kc):
    size_t hi_stride;   //   (r / bl) * hi_stride + (r % bl) * lo_stride + o * o_stride + kc * seg16
    size_t lo_stride;
    size_t o_stride;
    int bl;
};

__device__ __forceinline__ size_t addr(const Pat &p, int r, int o, int kc, int seg16)
{
    return (size_t) (r / p.bl) * p.hi_stride + (size_t) (r % p.bl) * p.lo_stride + (size_t) o * p.o_stride + (size_t) kc * seg16;
}

// one CTA per tile (persistent loop); N rows x seg16 float4 per row
template <int UNROLL>
__global__ void __launch_bounds__(1024) copy_tiles(const float4 *__restrict__ src, float4 *__restrict__ dst, Pat ps, Pat pd,
                                                  int nrows, int seg16, int ntk, int nouter)
{
    const int ntiles = ntk * nouter;
    const int lanes = seg16;                         // threads per row
    const int rpi = blockDim.x / lanes;              // rows per iteration
    const int c = threadIdx.x % lanes, t = threadIdx.x / lanes;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int o = tile / ntk, kc = tile - o * ntk;
        for (int r0 = 0; r0 < nrows; r0 += rpi * UNROLL) {
            float4 v[UNROLL];
            #pragma unroll
            for (int u = 0; u < UNROLL; u++) v[u] = __ldcs(src + addr(ps, r0 + u * rpi + t, o, kc, seg16) + c);
            #pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                float4 y = v[u]; y.x += 1.f;
                __stcs(dst + addr(pd, r0 + u * rpi + t, o, kc, seg16) + c, y);
            }
        }
    }
}

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 2048;
    const int nouter = argc > 2 ? atoi(argv[2]) : N;        // planes actually processed (timing sample)
    const size_t pitch_c = ((N / 2 + 1 + 15) / 16) * 16;    // complex per row
    const size_t pitch16 = pitch_c / 2;                     // float4 per row
    const size_t plane16 = (size_t) N * pitch16;
    const size_t total16 = (size_t) N * plane16;
    float4 *a, *b;
    if (cudaMalloc(&a, total16 * 16) || cudaMalloc(&b, total16 * 16)) { printf("alloc failed\n"); return 1; }
    cudaMemset(a, 0, total16 * 16); cudaMemset(b, 0, total16 * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("N=%d pitch_c=%zu mesh=%.2f GB nouter=%d\n", N, pitch_c, total16 * 16 / 1e9, nouter);

    struct Case { const char *name; Pat ps, pd; };
    auto run = [&](const char *name, Pat ps, Pat pd, int segB, int threads, int ctas_per_sm) {
        const int seg16 = segB / 16;
        const int ntk = (int) (pitch16 / seg16);
        const int grid = 148 * ctas_per_sm;
        float best = 1e30f;
        for (int it = 0; it < 3; it++) {
            cudaEventRecord(e0);
            copy_tiles<4><<<grid, threads>>>(a, b, ps, pd, N, seg16, ntk, nouter);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        const double bytes = 2.0 * nouter * (double) N * ntk * segB;
        printf("%-44s seg=%4dB thr=%4d x%d : %8.3f ms  %7.1f GB/s %s\n", name, segB, threads, ctas_per_sm, best, bytes / best / 1e6,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    // patterns.  "rowmajor": tile (o = plane, kc), rows at stride pitch (in-place style pass)
    Pat inplane = { pitch16, 0, plane16, 1 };
    // transposed destination [row][o][kz]: row stride = plane, o stride = pitch (current F2 / B1 writes)
    Pat transp = { plane16, 0, pitch16, 1 };
    for (int seg : { 64, 128, 256 }) {
        run("read in-plane, write in-plane", inplane, inplane, seg, 512, 2);
        run("read in-plane, write transposed(plane stride)", inplane, transp, seg, 512, 2);
        run("read transposed, write in-plane", transp, inplane, seg, 512, 2);
        run("read transposed, write transposed", transp, transp, seg, 512, 2);
    }
    // blocked k layout [r_hi][o][r_lo][kz]
    for (int bl : { 2, 4, 8, 16, 32, 64 }) {
        Pat blocked_w = { (size_t) N * bl * pitch16, pitch16, (size_t) bl * pitch16, bl };
        char nm[96]; snprintf(nm, sizeof nm, "read in-plane, write blocked bl=%d", bl);
        run(nm, inplane, blocked_w, 64, 512, 2);
        // the in-place pass over that layout: tile (o = other-axis index q, kc), rows r at stride bl*pitch:
        //   addr = (q / bl) * N*bl*pitch + r * bl*pitch + (q % bl) * pitch  -> emulate with o_stride = pitch for q%bl only (sample)
        Pat blocked_ip = { (size_t) bl * pitch16, 0, pitch16, 1 };      // q in [0, bl): exact; larger q aliases but same page behaviour
        snprintf(nm, sizeof nm, "in-place pass on blocked layout bl=%d", bl);
        const int save = nouter; (void) save;
        run(nm, blocked_ip, blocked_ip, 64, 512, 2);
    }
    // thread-count / occupancy sensitivity of the bad case
    run("read in-plane, write transposed", inplane, transp, 64, 1024, 1);
    run("read in-plane, write transposed", inplane, transp, 64, 256, 4);
    run("read in-plane, write transposed", inplane, transp, 64, 256, 8);
    run("read in-plane, write in-plane", inplane, inplane, 64, 256, 8);
    return 0;
}
