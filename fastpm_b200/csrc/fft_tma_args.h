// fastpm_b200 -- argument block of the TMA FFT tile pass (fft_tma.cu), shared with its caller (fft.cu)
#pragma once
#include "common.cuh"
#include "mesh.cuh"

struct TmaPassArgs {
    float2 *dst[FPM_MAX_RANKS];
    int rows_per_rank;
    size_t dst_estride, dst_ostride;
    int dst_ooffset;
    // staged slab transpose: rows owned by rank `self_rank` (>= 0) bypass the staging block and go to their final place
    int self_rank, self_ooffset;
    float2 *self_dst;
    size_t self_estride, self_ostride;
    int ntile_k, nouter;
    int conj;
    int chunk;                  // kz-adjacent tiles a CTA processes back to back
    int early;                  // 1: prefetch the next tile before the arithmetic of this one
    int outer0;
    const float2 *tw;           // [N] exp(-2 pi i t / N)
    FpmTransferSpec xfer;
    FpmKTables kt;
};

int fpm_fft_tma_pass(int n, const float2 *src, int pitch_c, int nouter, const TmaPassArgs &args, cudaStream_t st);
