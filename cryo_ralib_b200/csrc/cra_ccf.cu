// cra_ccf.cu -- stage 3+4 of the hot path: Crosrng_ms ring contraction fused with the
// inverse FFT and the peak search (EMAN2 Util::Crosrng_ms + the best-of loop of
// Util::multiref_polar_ali_2d; reference call site test_mref.py:200-201).
// Replaces cu_ccf_mult_m + cuFFT C2R + cu_max_idx_batch + cu_find_params
// (cuda/gpu_aln_noref.cu:1009-1143, :2198-2206, :1305-1346, :1393-1494) without ever
// materialising the particles x refs x shifts x 2 x (maxrin+2) CCF table in HBM.
//
// ccf_peak_kernel: one CTA = TM particle-shift rows x TN references.  Thread k owns
// angular frequency k and keeps, for every (row, ref) pair of the tile, the four
// real ring sums A=sum c1 d1, B=sum c2 d2, C=sum c1 d2, D=sum c2 d1 in registers
// (c = weighted reference spectrum, d = particle spectrum).  From them
//   q_k = (A+B) + i(D-C)   (straight,  ref * conj(img))
//   t_k = (A-B) - i(C+D)   (mirrored,  conj(ref) * conj(img))
// and the Hermitian-packed W = q + i t goes to shared memory, where one length-maxrin
// complex inverse FFT per pair (two register passes N1 x N2) yields q[m] + i t[m].
// The argmax over m (">=": last maximum wins, as the reference), the straight/mirror
// choice and the best-over-references rule are applied in registers/shuffles; only one
// (value, code) candidate per row x reference tile reaches HBM.
//
// finalize_kernel: one warp per particle scans its candidates in the reference's visit
// order (y outer, x inner, reference innermost, ">="), re-evaluates the 7 samples around
// the winning lag in double precision for prb1d, and emits the 6-tuple of
// multiref_polar_ali_2d.
#include "cra_common.cuh"
#include <math.h>

namespace {

constexpr int TM = 4;   // rows per CTA
constexpr int TN = 4;   // references per CTA
constexpr int NP = TM * TN;

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
__host__ __device__ constexpr int brevc(int i, int bits) { return bits == 0 ? 0 : ((i & 1) << (bits - 1)) | brevc(i >> 1, bits - 1); }

// In-register R-point DFT with kernel exp(+2 pi i n k / R) (inverse sign), R <= 32.
// Template recursion over the radix-2 stages keeps every register index a constant.
template <int R, int LEN>
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
                else if (tj == 8) v = make_float2(-b.y, b.x);
                else {
                    const float wr = tw_cos(tj), wi = tw_sin(tj);
                    v = make_float2(b.x * wr - b.y * wi, b.x * wi + b.y * wr);
                }
                x[i0] = make_float2(u.x + v.x, u.y + v.y);
                x[i1] = make_float2(u.x - v.x, u.y - v.y);
            }
        }
        FftStage<R, LEN * 2>::run(x);
    }
};
template <int R>
struct FftStage<R, 2 * R> { static __device__ __forceinline__ void run(float2 (&)[R]) {} };

template <int R>
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
    FftStage<R, 2>::run(x);
}

template <int LOG2N>
struct Shape {
    static constexpr int N = 1 << LOG2N;
    static constexpr int L1 = LOG2N / 2;
    static constexpr int L2 = LOG2N - L1;
    static constexpr int N1 = 1 << L1;
    static constexpr int N2 = 1 << L2;
    static constexpr int NT = (N / 2 < 32) ? 32 : N / 2;     // threads per CTA
    static constexpr int PS = N1 * (N2 + 1);                 // padded float2 stride of one pair
};

__device__ __forceinline__ bool better(float v, int m, float bv, int bm)
{   // ">=" scan order semantics: larger value wins, ties go to the later index
    return (v > bv) || (v == bv && m > bm);
}

template <int LOG2N>
__global__ void __launch_bounds__(Shape<LOG2N>::NT)
ccf_peak_kernel(const float* __restrict__ spec, int nrows, const float* __restrict__ refspec, int R,
                const CraRingTab* __restrict__ tab, const float2* __restrict__ twid,
                CraCand* __restrict__ cand, int ntile_n)
{
    using S = Shape<LOG2N>;
    constexpr int N = S::N, N1 = S::N1, N2 = S::N2, NT = S::NT, PS = S::PS;
    extern __shared__ __align__(16) float2 s_dyn[];
    float2* s_w = s_dyn;                 // NP * PS
    float2* s_tw = s_dyn + NP * PS;      // N : exp(+2 pi i j / N)
    __shared__ int s_off[CRA_MAX_RINGS];
    __shared__ int s_len[CRA_MAX_RINGS];
    __shared__ CraCand s_pair[NP];

    const int tid = threadIdx.x;
    const int tn = blockIdx.x % ntile_n;
    const int tm = blockIdx.x / ntile_n;
    const int nring = tab->nring;
    const int lcirc = tab->lcirc;
    for (int i = tid; i < nring; i += NT) { s_off[i] = tab->off[i]; s_len[i] = tab->len[i]; }
    for (int i = tid; i < N; i += NT) s_tw[i] = twid[i];
    __syncthreads();

    const float* drow[TM];
    const float* cref[TN];
#pragma unroll
    for (int m = 0; m < TM; ++m) { int r = tm * TM + m; if (r > nrows - 1) r = nrows - 1; drow[m] = spec + (size_t)r * lcirc; }
#pragma unroll
    for (int n = 0; n < TN; ++n) { int r = tn * TN + n; if (r > R - 1) r = R - 1; cref[n] = refspec + (size_t)r * lcirc; }

    const int k = tid;
    if (k < N / 2) {
        float A[TM][TN], B[TM][TN], Cc[TM][TN], D[TM][TN];
#pragma unroll
        for (int m = 0; m < TM; ++m)
#pragma unroll
            for (int n = 0; n < TN; ++n) { A[m][n] = 0.f; B[m][n] = 0.f; Cc[m][n] = 0.f; D[m][n] = 0.f; }

        for (int i = 0; i < nring; ++i) {
            const int len = s_len[i];
            if (k < (len >> 1)) {
                const int o = s_off[i] + 2 * k;
                float2 d[TM], c[TN];
#pragma unroll
                for (int m = 0; m < TM; ++m) d[m] = __ldg(reinterpret_cast<const float2*>(drow[m] + o));
#pragma unroll
                for (int n = 0; n < TN; ++n) c[n] = __ldg(reinterpret_cast<const float2*>(cref[n] + o));
                // thread 0: slot1 is the Nyquist term, which only full-length rings keep in place
                const float bsel = (k != 0 || len == N) ? 1.0f : 0.0f;
#pragma unroll
                for (int m = 0; m < TM; ++m)
#pragma unroll
                    for (int n = 0; n < TN; ++n) {
                        A[m][n] = fmaf(c[n].x, d[m].x, A[m][n]);
                        B[m][n] = fmaf(c[n].y * bsel, d[m].y, B[m][n]);
                        Cc[m][n] = fmaf(c[n].x, d[m].y, Cc[m][n]);
                        D[m][n] = fmaf(c[n].y, d[m].x, D[m][n]);
                    }
            } else if (k == (len >> 1) && len != N) {
                // Nyquist of a short ring: real term at frequency len/2 (Crosrng_ms q(numr3i+1))
                const int o = s_off[i] + 1;
#pragma unroll
                for (int m = 0; m < TM; ++m) {
                    const float dv = __ldg(drow[m] + o);
#pragma unroll
                    for (int n = 0; n < TN; ++n) A[m][n] = fmaf(__ldg(cref[n] + o), dv, A[m][n]);
                }
            }
        }
        // W = q + i t, Hermitian-extended to N points
#pragma unroll
        for (int m = 0; m < TM; ++m)
#pragma unroll
            for (int n = 0; n < TN; ++n) {
                float2* w = s_w + (m * TN + n) * PS;
                const float a = A[m][n], b = B[m][n], c = Cc[m][n], d = D[m][n];
                if (k == 0) {
                    w[0] = make_float2(a, a);
                    const int h = N / 2;
                    w[(h >> S::L2) * (N2 + 1) + (h & (N2 - 1))] = make_float2(b, b);
                } else {
                    const int kk = N - k;
                    w[(k >> S::L2) * (N2 + 1) + (k & (N2 - 1))] = make_float2(a + b + c + d, a - b + d - c);
                    w[(kk >> S::L2) * (N2 + 1) + (kk & (N2 - 1))] = make_float2(a + b - c - d, a - b + c - d);
                }
            }
    }
    __syncthreads();

    // pass 1: for each (pair, n2): N1-point DFT over n1 (stride N2), twiddle by w_N^(n2*k1)
    for (int item = tid; item < NP * N2; item += NT) {
        const int pair = item / N2, n2 = item % N2;
        float2* w = s_w + pair * PS + n2;
        float2 x[N1];
#pragma unroll
        for (int j = 0; j < N1; ++j) x[j] = w[j * (N2 + 1)];
        fft_reg<N1>(x);
#pragma unroll
        for (int j = 0; j < N1; ++j) {
            const float2 t = s_tw[(n2 * j) & (N - 1)];
            w[j * (N2 + 1)] = make_float2(x[j].x * t.x - x[j].y * t.y, x[j].x * t.y + x[j].y * t.x);
        }
    }
    __syncthreads();

    // pass 2: for each (pair, k1): N2-point DFT over n2 -> X[k1 + N1*k2]; argmax over lags
    for (int item = tid; item < NP * N1; item += NT) {
        const int pair = item / N1, k1 = item % N1;
        const float2* w = s_w + pair * PS + k1 * (N2 + 1);
        float2 x[N2];
#pragma unroll
        for (int j = 0; j < N2; ++j) x[j] = w[j];
        fft_reg<N2>(x);
        float bq = -INFINITY, bt = -INFINITY; int mq = -1, mt = -1;
#pragma unroll
        for (int j = 0; j < N2; ++j) {
            const int m = k1 + N1 * j;
            if (x[j].x >= bq) { bq = x[j].x; mq = m; }
            if (x[j].y >= bt) { bt = x[j].y; mt = m; }
        }
#pragma unroll
        for (int o = N1 >> 1; o > 0; o >>= 1) {
            float oq = __shfl_xor_sync(0xffffffffu, bq, o); int omq = __shfl_xor_sync(0xffffffffu, mq, o);
            float ot = __shfl_xor_sync(0xffffffffu, bt, o); int omt = __shfl_xor_sync(0xffffffffu, mt, o);
            if (better(oq, omq, bq, mq)) { bq = oq; mq = omq; }
            if (better(ot, omt, bt, mt)) { bt = ot; mt = omt; }
        }
        if (k1 == 0) {
            const int m = pair / TN, n = pair % TN;
            const int row = tm * TM + m, ref = tn * TN + n;
            const float sc = 1.0f / (float)N;
            const float qn = bq * sc, qm = bt * sc;
            CraCand cd;
            if (row < nrows && ref < R) {
                if (qn >= qm) { cd.v = qn; cd.code = ref * 8192 + (mq + 1); }
                else          { cd.v = qm; cd.code = ref * 8192 + 4096 + (mt + 1); }
            } else { cd.v = -INFINITY; cd.code = -1; }
            s_pair[pair] = cd;
        }
    }
    __syncthreads();
    if (tid < TM) {
        const int row = tm * TM + tid;
        if (row < nrows) {
            CraCand best; best.v = -INFINITY; best.code = -1;
#pragma unroll
            for (int n = 0; n < TN; ++n) {
                const CraCand c = s_pair[tid * TN + n];
                if (c.code >= 0 && c.v >= best.v) best = c;
            }
            cand[(size_t)row * ntile_n + tn] = best;
        }
    }
}

// Ring sums of one (row, ref) pair at frequency k (0 < k < N/2), incl. short-ring Nyquist.
__device__ void pair_freq(const float* __restrict__ d, const float* __restrict__ c, const CraRingTab* tab,
                          int k, int N, float& zq_r, float& zq_i, float& zt_r, float& zt_i)
{
    float A = 0.f, B = 0.f, C = 0.f, D = 0.f;
    for (int i = 0; i < tab->nring; ++i) {
        const int len = tab->len[i];
        if (k < (len >> 1)) {
            const int o = tab->off[i] + 2 * k;
            const float c1 = c[o], c2 = c[o + 1], d1 = d[o], d2 = d[o + 1];
            A = fmaf(c1, d1, A); B = fmaf(c2, d2, B); C = fmaf(c1, d2, C); D = fmaf(c2, d1, D);
        } else if (k == (len >> 1) && len != N) {
            const int o = tab->off[i] + 1;
            A = fmaf(c[o], d[o], A);
        }
    }
    zq_r = A + B; zq_i = D - C; zt_r = A - B; zt_i = -C - D;
}
__device__ void pair_dc_nyq(const float* __restrict__ d, const float* __restrict__ c, const CraRingTab* tab,
                            int N, float& dc, float& nyq)
{
    float a = 0.f, b = 0.f;
    for (int i = 0; i < tab->nring; ++i) {
        const int o = tab->off[i];
        a = fmaf(c[o], d[o], a);
        if (tab->len[i] == N) b = fmaf(c[o + 1], d[o + 1], b);
    }
    dc = a; nyq = b;
}

__global__ void __launch_bounds__(128)
finalize_kernel(const float* __restrict__ spec, const float* __restrict__ refspec, int R,
                const CraRingTab* __restrict__ tab, const CraCand* __restrict__ cand, int ntile_n,
                CraRowMap map, CraResult* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= map.np) return;
    const int r0 = map.row_start[p], r1 = map.row_start[p + 1];
    const int ncand = (r1 - r0) * ntile_n;
    float bv = -INFINITY; int bc = -1, bcode = -1;
    const CraCand* cd = cand + (size_t)r0 * ntile_n;
    for (int c = lane; c < ncand; c += 32) {
        const CraCand x = cd[c];
        if (x.code >= 0 && better(x.v, c, bv, bc)) { bv = x.v; bc = c; bcode = x.code; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oc = __shfl_xor_sync(0xffffffffu, bc, o);
        int ocode = __shfl_xor_sync(0xffffffffu, bcode, o);
        if (oc >= 0 && better(ov, oc, bv, bc)) { bv = ov; bc = oc; bcode = ocode; }
    }
    CraResult res;
    if (bc < 0) {   // empty window: cannot happen for valid requests
        if (lane == 0) { res.ang = 0; res.sxs = 0; res.sys = 0; res.mirror = 0; res.iref = 0; res.peak = -1.0e23f; res.sx = 0; res.sy = 0; out[p] = res; }
        return;
    }
    const int row = r0 + bc / ntile_n;
    const int iref = bcode / 8192;
    const int mirror = (bcode >> 12) & 1;
    const int jtot = bcode & 4095;              // 1-based lag of the maximum
    const int N = tab->maxrin;
    const float* d = spec + (size_t)row * tab->lcirc;
    const float* c = refspec + (size_t)iref * tab->lcirc;

    double t7[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int k = 1 + lane; k < N / 2; k += 32) {
        float qr, qi, tr, ti;
        pair_freq(d, c, tab, k, N, qr, qi, tr, ti);
        const double zr = mirror ? tr : qr, zi = mirror ? ti : qi;
#pragma unroll
        for (int s = 0; s < 7; ++s) {
            const int m = (jtot - 1 + s - 3 + N) % N;                 // 0-based lag
            const int ph = (int)(((long long)k * m) % N);
            double sn, cs; sincospi(2.0 * (double)ph / (double)N, &sn, &cs);
            t7[s] += 2.0 * (zr * cs - zi * sn);
        }
    }
#pragma unroll
    for (int s = 0; s < 7; ++s)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t7[s] += __shfl_xor_sync(0xffffffffu, t7[s], o);
    if (lane == 0) {
        float dc, nyq; pair_dc_nyq(d, c, tab, N, dc, nyq);
#pragma unroll
        for (int s = 0; s < 7; ++s) {
            const int m = (jtot - 1 + s - 3 + N) % N;
            t7[s] = (t7[s] + dc + ((m & 1) ? -(double)nyq : (double)nyq)) / (double)N;
        }
        const double c2 = 49. * t7[0] + 6. * t7[1] - 21. * t7[2] - 32. * t7[3] - 27. * t7[4] - 6. * t7[5] + 31. * t7[6];
        const double c3 = 5. * t7[0] - 3. * t7[2] - 4. * t7[3] - 3. * t7[4] + 5. * t7[6];
        float pos = 0.0f;
        if (c3 != 0.0) pos = (float)((c2 / (2.0 * c3)) - 4);
        const float tot = (float)jtot + pos;
        const float ang = fmodf(((tot - 1.0f) / N + 1.0f) * 360.0f, 360.0f);
        const int li = row - r0;
        const int4 w = map.win[p];
        const int wx = w.x + w.y + 1;
        const float iy = (li / wx - w.z) * map.step, ix = (li % wx - w.x) * map.step;
        const float sx = -ix, sy = -iy;
        const float co = (float)cos((double)ang * 3.14159265358979323846 / 180.0);
        const float so = (float)(-sin((double)ang * 3.14159265358979323846 / 180.0));
        res.ang = ang; res.sxs = sx * co - sy * so; res.sys = sx * so + sy * co;
        res.mirror = mirror; res.iref = iref; res.peak = bv; res.sx = sx; res.sy = sy;
        out[p] = res;
    }
}

// Test entry: full q/t curves of one pair by direct evaluation (block of N threads).
__global__ void ccf_curves_kernel(const float* __restrict__ d, const float* __restrict__ c,
                                  const CraRingTab* __restrict__ tab, float* __restrict__ q, float* __restrict__ t)
{
    extern __shared__ float4 s_z[];   // N/2 entries: (qr, qi, tr, ti)
    const int N = tab->maxrin;
    __shared__ float s_dc, s_nyq;
    for (int k = threadIdx.x; k < N / 2; k += blockDim.x) {
        float4 z = make_float4(0, 0, 0, 0);
        if (k > 0) pair_freq(d, c, tab, k, N, z.x, z.y, z.z, z.w);
        s_z[k] = z;
    }
    if (threadIdx.x == 0) pair_dc_nyq(d, c, tab, N, s_dc, s_nyq);
    __syncthreads();
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
        double aq = 0, at = 0;
        for (int k = 1; k < N / 2; ++k) {
            const int ph = (int)(((long long)k * m) % N);
            double sn, cs; sincospi(2.0 * (double)ph / (double)N, &sn, &cs);
            const float4 z = s_z[k];
            aq += 2.0 * (z.x * cs - z.y * sn);
            at += 2.0 * (z.z * cs - z.w * sn);
        }
        const double ny = (m & 1) ? -(double)s_nyq : (double)s_nyq;
        q[m] = (float)((aq + s_dc + ny) / N);
        t[m] = (float)((at + s_dc + ny) / N);
    }
}

template <int LOG2N>
int launch_ccf_t(const float* spec, int nrows, const float* refspec, int R, const CraRingTab* tab,
                 const float2* twid, CraCand* cand, int ntile_n, cudaStream_t st)
{
    using S = Shape<LOG2N>;
    const size_t smem = ((size_t)NP * S::PS + S::N) * sizeof(float2);
    static bool configured = false;
    if (!configured) {
        CRA_CUDA(cudaFuncSetAttribute(ccf_peak_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const long ntile_m = (nrows + TM - 1) / TM;
    const long nblk = ntile_m * ntile_n;
    if (nblk <= 0) return 0;
    if (nblk > 2147483647L) { cra_set_error("ccf grid too large; lower row_batch"); return 1; }
    ccf_peak_kernel<LOG2N><<<(unsigned)nblk, S::NT, smem, st>>>(spec, nrows, refspec, R, tab, twid, cand, ntile_n);
    CRA_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int cra_ccf_tile_n() { return TN; }

int cra_launch_ccf(const float* spec, int nrows, const float* refspec, int R, const CraRingTab* tab,
                   const CraRingTab& htab, const float2* twid, CraCand* cand, int ntile_n, cudaStream_t st)
{
    switch (htab.log2n) {
        case 5:  return launch_ccf_t<5>(spec, nrows, refspec, R, tab, twid, cand, ntile_n, st);
        case 6:  return launch_ccf_t<6>(spec, nrows, refspec, R, tab, twid, cand, ntile_n, st);
        case 7:  return launch_ccf_t<7>(spec, nrows, refspec, R, tab, twid, cand, ntile_n, st);
        case 8:  return launch_ccf_t<8>(spec, nrows, refspec, R, tab, twid, cand, ntile_n, st);
        case 9:  return launch_ccf_t<9>(spec, nrows, refspec, R, tab, twid, cand, ntile_n, st);
        case 10: return launch_ccf_t<10>(spec, nrows, refspec, R, tab, twid, cand, ntile_n, st);
        default: cra_set_error("maxrin must be a power of two in [32, 1024]"); return 1;
    }
}

int cra_launch_finalize(const float* spec, const float* refspec, int R, const CraRingTab* tab, const CraRingTab& htab,
                        const CraCand* cand, int ntile_n, CraRowMap map, CraResult* out, cudaStream_t st)
{
    (void)htab;
    if (map.np <= 0) return 0;
    const int wpb = 4;
    finalize_kernel<<<(map.np + wpb - 1) / wpb, wpb * 32, 0, st>>>(spec, refspec, R, tab, cand, ntile_n, map, out);
    CRA_CUDA(cudaGetLastError());
    return 0;
}

int cra_launch_ccf_curves(const float* spec_row, const float* refspec_row, const CraRingTab* tab, const CraRingTab& htab,
                          float* q, float* t, cudaStream_t st)
{
    ccf_curves_kernel<<<1, 256, (htab.maxrin / 2) * sizeof(float4), st>>>(spec_row, refspec_row, tab, q, t);
    CRA_CUDA(cudaGetLastError());
    return 0;
}
