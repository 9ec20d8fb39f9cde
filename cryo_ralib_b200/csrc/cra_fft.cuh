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
                float2 u = x[i0], b = x[i1], v;
                if (tj == 0) v = b;
                else if (tj == 8) v = (SGN > 0) ? make_float2(-b.y, b.x) : make_float2(b.y, -b.x);
                else {
                    const float wr = tw_cos(tj), wi = SGN * tw_sin(tj);
                    v = make_float2(b.x * wr - b.y * wi, b.x * wi + b.y * wr);
                }
                x[i0] = make_float2(u.x + v.x, u.y + v.y);
                x[i1] = make_float2(u.x - v.x, u.y - v.y);
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
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

}  // namespace crafft
