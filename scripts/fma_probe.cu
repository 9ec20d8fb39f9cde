// Micro-benchmark: FP32 FMA rate of a register-resident 4x4 complex-MAC tile (64 accumulators per
// thread, the CCF contraction's inner step) versus resident warps per SM, scalar FFMA vs packed FFMA2.
#include <cstdio>
#include <cuda_runtime.h>
template <int PACKED>
__global__ void tile_kernel(float* out, int iters, float seed)
{
    float2 ab[4][4], cd[4][4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) { ab[m][n] = make_float2(0.f, 0.f); cd[m][n] = make_float2(0.f, 0.f); }
    float2 d[4], c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { d[i] = make_float2(seed + i + threadIdx.x, seed - i); c[i] = make_float2(seed * i, 1.0f + i); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                if (PACKED) {
                    ab[m][n] = __ffma2_rn(c[n], d[m], ab[m][n]);
                    cd[m][n] = __ffma2_rn(c[n], make_float2(d[m].y, d[m].x), cd[m][n]);
                } else {
                    ab[m][n].x = fmaf(c[n].x, d[m].x, ab[m][n].x); ab[m][n].y = fmaf(c[n].y, d[m].y, ab[m][n].y);
                    cd[m][n].x = fmaf(c[n].x, d[m].y, cd[m][n].x); cd[m][n].y = fmaf(c[n].y, d[m].x, cd[m][n].y);
                }
            }
#pragma unroll
        for (int i = 0; i < 4; ++i) { d[i].x += 1e-3f; c[i].y -= 1e-3f; }   // operands change every step
    }
    float s = 0.f;
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) s += ab[m][n].x + ab[m][n].y + cd[m][n].x + cd[m][n].y;
    if (s == 123.456f) out[0] = s;
}
int main()
{
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int packed = 0; packed < 2; ++packed)
        for (int wps = 4; wps <= 32; wps += (wps < 16 ? 4 : 8)) {     // resident warps per SM
            // 96-thread CTAs (3 warps); grid sized to exactly fill: wps/3 CTAs per SM via smem limit trick is
            // unnecessary: use one CTA per SM with wps warps
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (packed) tile_kernel<1><<<148, wps * 32>>>(d, iters, 1.0f + rep);
                else tile_kernel<0><<<148, wps * 32>>>(d, iters, 1.0f + rep);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            double flops = 2.0 * 64 * (double)iters * wps * 32 * 148;
            printf("%s warps/SM %2d : %.2f TFLOP/s (%.1f%% of 70.5)\n", packed ? "FFMA2" : "FFMA ", wps, flops / best / 1e9, flops / best / 1e9 / 70.5 * 100);
        }
    return 0;
}
