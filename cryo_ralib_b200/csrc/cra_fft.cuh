// cra_fft.cuh -- in-register radix-2 DFTs of 2..32 points with compile-time indices and
// immediate twiddles, shared by the ring FFTs (forward) and the CCF inverse FFT.
#pragma once
#include <cuda_runtime.h>

namespace crafft {

__host__ __device__ constexpr float tw_cos(int j)   // cos(2 pi j / 32), j < 16
{
    return j == 0 ? 1.0f : j == 1 ? 9.807852804e-01f : j == 2 ? 9.238795325e-01f : j == 3 ? 8.314696123e-01f
         : j == 4 ? 7.071067812e-01f : j == 5 ? 5.555702330e-01f : j == 6 ? 3.826834324e-01f : j == 7 ? 1.950903220e-01f
         : j == 8 ? 0.0f : j == 9 ? -1.950903220e-01f : j == 10 ? -3.826834324e-01f : j == 11 ? -5.555702330e-01f
         : j == 12 ? -7.071067812e-01f : j == 13 ? -8.314696123e-01f : j == 14 ? -9.238795325e-01f : -9.807852804e-01f;
}
__host__ __device__ constexpr float tw_sin(int j)   // sin(2 pi j / 32), j < 16
{
    return j == 0 ? 0.0f : j == 1 ? 1.950903220e-01f : j == 2 ? 3.826834324e-01f : j == 3 ? 5.555702330e-01f
         : j == 4 ? 7.071067812e-01f : j == 5 ? 8.314696123e-01f : j == 6 ? 9.238795325e-01f : j == 7 ? 9.807852804e-01f
         : j == 8 ? 1.0f : j == 9 ? 9.807852804e-01f : j == 10 ? 9.238795325e-01f : j == 11 ? 8.314696123e-01f
         : j == 12 ? 7.071067812e-01f : j == 13 ? 5.555702330e-01f : j == 14 ? 3.826834324e-01f : 1.950903220e-01f;
}
__host__ __device__ constexpr int ilog2c(int n) { return n <= 1 ? 0 : 1 + ilog2c(n >> 1); }

// Packed FP32 helpers: sm_100 FFMA2 / FADD2 / FMUL2 work on (lo, hi) register pairs, and ptxas
// folds swapped (.LO_HI) and broadcast (.F32) operands into the instruction, so a complex
// add is one instruction and a complex multiply two.
__device__ __forceinline__ float2 swp(float2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
// a * (wr + i wi)
__device__ __forceinline__ float2 cmul_w(float2 a, float wr, float wi)
{
    return __ffma2_rn(swp(a), make_float2(-wi, wi), __fmul2_rn(a, make_float2(wr, wr)));
}

// SGN = +1: kernel exp(+2 pi i n k / R) (inverse);  SGN = -1: exp(-2 pi i n k / R) (forward)
template <int R, int LEN, int SGN>
struct FftStage {
    static __device__ __forceinline__ void run(float2 (&x)[R])
    {
        constexpr int HALF = LEN / 2;
#pragma unroll
        for (int g = 0; g < R / LEN; ++g) {
#pragma unroll
            for (int k = 0; k < HALF; ++k) {
                constexpr int TS = 32 / LEN;
                const int tj = k * TS;
                const int i0 = g * LEN + k, i1 = i0 + HALF;
                const float2 u = x[i0], b = x[i1];
                if (tj == 0) {
                    x[i0] = cadd(u, b);
                    x[i1] = csub(u, b);
                } else if (tj == 8) {
                    // v = (+-i) b = SGN * (-b.y, b.x);  u +- v folded into one packed FMA each
                    const float2 sb = swp(b);
                    x[i0] = __ffma2_rn(sb, make_float2(-(float)SGN, (float)SGN), u);
                    x[i1] = __ffma2_rn(sb, make_float2((float)SGN, -(float)SGN), u);
                } else {
                    const float2 v = cmul_w(b, tw_cos(tj), SGN * tw_sin(tj));
                    x[i0] = cadd(u, v);
                    x[i1] = csub(u, v);
                }
            }
        }
        FftStage<R, LEN * 2, SGN>::run(x);
    }
};
template <int R, int SGN>
struct FftStage<R, 2 * R, SGN> { static __device__ __forceinline__ void run(float2 (&)[R]) {} };

template <int R, int SGN>
__device__ __forceinline__ void fft_reg(float2 (&x)[R])
{
    constexpr int LG = ilog2c(R);
#pragma unroll
    for (int i = 0; i < R; ++i) {
        int j = 0;
#pragma unroll
        for (int b = 0; b < LG; ++b) j |= ((i >> b) & 1) << (LG - 1 - b);
        if (i < j) { float2 t = x[i]; x[i] = x[j]; x[j] = t; }
    }
    FftStage<R, 2, SGN>::run(x);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return __ffma2_rn(swp(a), make_float2(-b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}

}  // namespace crafft
