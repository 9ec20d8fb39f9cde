// cra_ccf.cu -- stage 3+4 of the hot path: Crosrng_ms ring contraction fused with the
// inverse FFT and the peak search (EMAN2 Util::Crosrng_ms + the best-of loop of
// Util::multiref_polar_ali_2d; reference call site test_mref.py:200-201).
// Replaces cu_ccf_mult_m + cuFFT C2R + cu_max_idx_batch + cu_find_params
// (cuda/gpu_aln_noref.cu:1009-1143, :2198-2206, :1305-1346, :1393-1494) without ever
// materialising the particles x refs x shifts x 2 x (maxrin+2) CCF table in HBM.
//
// ccf_peak_kernel: one CTA = 4 particle-shift rows x 4 references (16 pairs).  A thread owns
// one angular frequency k (two for the upper quarter of the band, where fewer rings reach, so
// the warps of a CTA carry equal ring counts) and keeps, for every pair, the four real ring sums
//   A = sum c.x d.x, B = sum c.y d.y, C = sum c.x d.y, D = sum c.y d.x
// in registers (c = weighted reference spectrum, d = particle spectrum, both in the device
// layout of cra_common.cuh: 4 rows / 4 refs interleaved, so each ring step is four 128-bit
// loads, prefetched one ring ahead).  From them
//   q_k = (A+B) + i(D-C)   (straight,  ref * conj(img))
//   t_k = (A-B) - i(C+D)   (mirrored,  conj(ref) * conj(img))
// and the Hermitian-extended W = q + i t goes to shared memory, where one length-maxrin
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
#include "cra_fft.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>

namespace {

constexpr int RG = 1;        // row groups (of 4 rows) per CTA: each thread keeps RG x 4 x TNT pairs in registers
constexpr int RH = 1;        // thread sets per CTA, each owning TN/RH references of the tile (more warps per SM)
constexpr int TNT = 4 / RH;  // references per thread
constexpr int TM = 4 * RG;   // rows per CTA
constexpr int TN = 4;        // references per CTA
constexpr int NP = TM * TN;

// ring table of the active configuration (uniform-datapath reads in the hot loop)
__constant__ int c_coff[CRA_MAX_RINGS];
__constant__ int c_half[CRA_MAX_RINGS];   // len/2

using crafft::fft_reg;

template <int LOG2N>
struct Shape {
    static constexpr int N = 1 << LOG2N;
    static constexpr int L1 = LOG2N / 2;
    static constexpr int L2 = LOG2N - L1;
    static constexpr int N1 = 1 << L1;
    static constexpr int N2 = 1 << L2;
    static constexpr int NQ = N / 4;                          // threads owning one frequency
    static constexpr int NW = 3 * N / 8;                      // working threads (upper quarter: two frequencies each)
    static constexpr int NT1 = ((NW + 31) / 32) * 32 < 32 ? 32 : ((NW + 31) / 32) * 32;   // threads of one set
    static constexpr int NT = NT1 * RH;
    static constexpr int PS = N1 * (N2 + 1);                  // padded float2 stride of one pair
    // sub-tiles per CTA (row groups x reference groups), bounded by shared memory for the W buffers
    // (measured on B200: 2x2 sub-tiles relying on L1 sharing alone are slower than 1x1 -- 26.3 vs
    //  19.6 ms per 1e7 alignments; a cp.async.bulk/mbarrier staged 2x2 variant was also correct but
    //  slower, 37 ms -- see DESIGN.md 3.2.  Operand reuse is raised in registers instead: RG = 2.)
    static constexpr int SUBM = 1;
    static constexpr int SUBN = 1;
    static constexpr int NSUB = SUBM * SUBN;
};

__device__ __forceinline__ bool better(float v, int m, float bv, int bm)
{   // ">=" scan order semantics: larger value wins, ties go to the later index
    return (v > bv) || (v == bv && m > bm);
}

// ring sums as packed pairs: ab = (A, B) = sum c (.) d,  cd = (C, D) = sum c (.) swap(d)
struct Acc { float2 ab[TM][TNT], cd[TM][TNT]; };
constexpr int NCP = (TNT + 1) / 2;   // reference planes (float4 = 2 refs) a thread loads

__device__ __forceinline__ void ring_fma(Acc& a, const float4 (&d)[2 * RG], const float4 (&c)[NCP])
{
    float2 dv[TM], cv[TNT];
#pragma unroll
    for (int g = 0; g < RG; ++g) {
        dv[4 * g + 0] = make_float2(d[2 * g].x, d[2 * g].y); dv[4 * g + 1] = make_float2(d[2 * g].z, d[2 * g].w);
        dv[4 * g + 2] = make_float2(d[2 * g + 1].x, d[2 * g + 1].y); dv[4 * g + 3] = make_float2(d[2 * g + 1].z, d[2 * g + 1].w);
    }
#pragma unroll
    for (int g = 0; g < NCP; ++g) {
        cv[2 * g] = make_float2(c[g].x, c[g].y);
        if (2 * g + 1 < TNT) cv[2 * g + 1] = make_float2(c[g].z, c[g].w);
    }
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
        for (int n = 0; n < TNT; ++n) {
            // scalar FFMA on purpose: scripts/fma_probe.cu measures 53 TFLOP/s for this tile with FFMA
            // and 50 with FFMA2 on B200 -- packed math only pays where issue slots are the limit
            a.ab[m][n].x = fmaf(cv[n].x, dv[m].x, a.ab[m][n].x);   // A
            a.ab[m][n].y = fmaf(cv[n].y, dv[m].y, a.ab[m][n].y);   // B
            a.cd[m][n].x = fmaf(cv[n].x, dv[m].y, a.cd[m][n].x);   // C
            a.cd[m][n].y = fmaf(cv[n].y, dv[m].x, a.cd[m][n].y);   // D
        }
}

// Ring sums of all 16 pairs at frequency k, then W[k] and W[N-k] to shared memory.
// dq / cq point at element 0 of the first row group / the ref group (float4 = two interleaved
// float2); gstride is the float4 distance between consecutive row groups.
template <int LOG2N>
__device__ __forceinline__ void contract_freq(int k, int rh, int nring, const float4* __restrict__ dq, size_t gstride,
                                              const float4* __restrict__ cq, float2* __restrict__ s_w)
{
    using S = Shape<LOG2N>;
    constexpr int N = S::N, N2 = S::N2, PS = S::PS;
    Acc a;
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
        for (int n = 0; n < TNT; ++n) { a.ab[m][n] = make_float2(0.f, 0.f); a.cd[m][n] = make_float2(0.f, 0.f); }

    // rings in descending length: stop at the first ring that no longer reaches frequency k.
    // Operands are prefetched two rings ahead through three register buffers.
// timing experiments only (never defined in a shipped build): drop the operand loads
#ifdef CRA_EXP_NOROW
#define CRA_EXP_ROWLOAD(ptr) make_float4(1.0f + (float)e_, 0.5f, 0.25f + (float)p_, 2.0f)
#else
#define CRA_EXP_ROWLOAD(ptr) __ldg(ptr)
#endif
#ifdef CRA_EXP_NOREF
#define CRA_EXP_REFLOAD(ptr, alt) (alt)
#else
#define CRA_EXP_REFLOAD(ptr, alt) __ldg(ptr)
#endif
#define CRA_HAS(j) ((j) >= 0 && k < c_half[(j) < 0 ? 0 : (j)])
#define CRA_LOAD(D, C, j)                                                                        \
    { const int e_ = 2 * c_coff[j] + k, p_ = c_half[j] + 1;                                      \
      _Pragma("unroll") for (int g_ = 0; g_ < RG; ++g_) {                                        \
          D[2 * g_] = CRA_EXP_ROWLOAD(dq + g_ * gstride + e_); D[2 * g_ + 1] = CRA_EXP_ROWLOAD(dq + g_ * gstride + e_ + p_); } \
      _Pragma("unroll") for (int g_ = 0; g_ < NCP; ++g_)                                          \
          C[g_] = CRA_EXP_REFLOAD(cq + e_ + (RH == 1 ? g_ : rh) * p_, D[g_]); }
    int i = nring - 1;
    float4 d0[2 * RG], c0[NCP], d1[2 * RG], c1[NCP], d2[2 * RG], c2[NCP];
    CRA_LOAD(d0, c0, i);
    bool h1 = CRA_HAS(i - 1);
    if (h1) CRA_LOAD(d1, c1, i - 1);
    for (;;) {
        const bool h2 = h1 && CRA_HAS(i - 2);
        if (h2) CRA_LOAD(d2, c2, i - 2);
        ring_fma(a, d0, c0);
        if (!h1) break;
        const bool h3 = h2 && CRA_HAS(i - 3);
        if (h3) CRA_LOAD(d0, c0, i - 3);
        ring_fma(a, d1, c1);
        if (!h2) break;
        const bool h4 = h3 && CRA_HAS(i - 4);
        if (h4) CRA_LOAD(d1, c1, i - 4);
        ring_fma(a, d2, c2);
        if (!h3) break;
        h1 = h4;
        i -= 3;
    }
#undef CRA_HAS
#undef CRA_LOAD
    const int kk = (N - k) & (N - 1);
    const int i0 = (k >> S::L2) * (N2 + 1) + (k & (N2 - 1));
    const int i1 = (kk >> S::L2) * (N2 + 1) + (kk & (N2 - 1));
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
        for (int n = 0; n < TNT; ++n) {
            float2* w = s_w + (m * TN + rh * TNT + n) * PS;
            const float2 ab = a.ab[m][n], cd = a.cd[m][n];
            // s = (A+B, A-B), t = (C+D, D-C);  W[k] = s + t,  W[N-k] = s - t
            const float2 sv = __ffma2_rn(make_float2(ab.y, ab.y), make_float2(1.0f, -1.0f), make_float2(ab.x, ab.x));
            const float2 tv = __ffma2_rn(make_float2(cd.x, cd.x), make_float2(1.0f, -1.0f), make_float2(cd.y, cd.y));
            w[i0] = crafft::cadd(sv, tv);
            if (k != 0) w[i1] = crafft::csub(sv, tv);
        }
}

// A CTA runs SUBM x SUBN independent 16-pair sub-tiles (each NT threads) side by side: 2 row groups
// x 2 reference groups.  They walk the rings in step, so the row spectra fetched by one sub-tile are
// L1 hits for its neighbour and the L2 -> SM traffic per pair halves (the kernel is L2-bandwidth
// bound otherwise: 188 KB of operands per 16 pairs).
template <int LOG2N>
__global__ void __launch_bounds__(Shape<LOG2N>::NT * Shape<LOG2N>::NSUB)
ccf_peak_kernel(const float4* __restrict__ spec, int nrows, const float4* __restrict__ refspec, int R,
                int nring, int nc, const float2* __restrict__ twid, CraCand* __restrict__ cand, int ntile_n,
                int ntile_m, int ncta_n)
{
    using S = Shape<LOG2N>;
    constexpr int N = S::N, N1 = S::N1, N2 = S::N2, NT = S::NT, NT1 = S::NT1, PS = S::PS, NQ = S::NQ, NW = S::NW;
    constexpr int SUBM = S::SUBM, SUBN = S::SUBN, NSUB = S::NSUB;
    extern __shared__ __align__(16) float2 s_dyn[];
    const int sub = threadIdx.x / NT;
    float2* s_w = s_dyn + sub * (NP * PS);          // NP * PS per sub-tile
    float2* s_tw = s_dyn + NSUB * (NP * PS);        // N : s_tw[j*N2 + n2] = exp(+2 pi i n2 j / N)
    __shared__ CraCand s_pair_all[NSUB][NP];
    CraCand* s_pair = s_pair_all[sub];

    const int tid = threadIdx.x - sub * NT;
    // reference groups fastest so that concurrently resident CTAs share row groups in L2
    int tn = (blockIdx.x % ncta_n) * SUBN + (sub % SUBN);
    int tm = (blockIdx.x / ncta_n) * SUBM + (sub / SUBN);
    const bool live = (tn < ntile_n) && (tm < ntile_m);
    if (tn >= ntile_n) tn = ntile_n - 1;            // idle sub-tiles still take part in the barriers
    if (tm >= ntile_m) tm = ntile_m - 1;
    for (int i = threadIdx.x; i < N; i += NT * NSUB) s_tw[i] = twid[i];

    // row groups tm*RG .. tm*RG+RG-1 (the last CTA may run past the batch: reuse the last group)
    const int ngroups = (nrows + 3) >> 2;
    const size_t gstride = (tm * RG + RG - 1 < ngroups) ? (size_t)nc * 2 : 0;
    const float4* dq = spec + (size_t)tm * RG * nc * 2;   // nc float2x4 = nc*2 float4 per group
    const float4* cq = refspec + (size_t)tn * nc * 2;

    {
        const int rh = tid / NT1, t1 = tid - rh * NT1;      // thread set (reference half) and its frequency slot
        if (t1 < NW) {
            contract_freq<LOG2N>(t1, rh, nring, dq, gstride, cq, s_w);
            if (t1 >= NQ) contract_freq<LOG2N>(t1 + N / 8, rh, nring, dq, gstride, cq, s_w);
        }
    }
    // Real-valued ring elements.  (i) frequency N/2: only full-length rings reach it; one pair per
    // lane of the last warp.  (ii) the Nyquist term of every shorter ring (frequency len/2 < N/2,
    // Crosrng_ms q(numr3i+1)): summed per length class by the lanes of warp 1 % nwarps, the least
    // loaded warp, and folded into W after the barrier.
    __shared__ float s_nyq_all[NSUB][NP][8];
    float (*s_nyq)[8] = s_nyq_all[sub];
    const int nl = tid - ((NT >= 64) ? 32 : 0);
    const bool nyq_lane = (nl >= 0 && nl < 32);
    {
        const float2* d2 = reinterpret_cast<const float2*>(dq);
        const float2* c2 = reinterpret_cast<const float2*>(cq);
        const int l = tid - (NT - 32);
        // Both loops issue their loads in batches of 8 rings: done one ring at a time they are a serial
        // chain of L2 round trips (16-20 x ~600 clk) on which the whole CTA then waits at the barrier.
        if (l >= 0 && l < 32) {
            int nfull = 0;
            for (int i = nring - 1; i >= 0 && c_half[i] == N / 2; --i) ++nfull;
            for (int pair = l; pair < NP; pair += 32) {
                const int m = pair / TN, n = pair % TN;
                const float2* dm = d2 + (size_t)(m >> 2) * gstride * 2;
                float a = 0.f;
                for (int base = 0; base < nfull; base += 8) {
                    float cvv[8], dvv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const bool ok = base + u < nfull;
                        const int i = ok ? nring - 1 - (base + u) : nring - 1;
                        cvv[u] = ok ? __ldg(c2 + cra_spec_idx(c_coff[i], N / 2, n, N / 2)).x : 0.f;
                        dvv[u] = ok ? __ldg(dm + cra_spec_idx(c_coff[i], N / 2, m, N / 2)).x : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) a = fmaf(cvv[u], dvv[u], a);
                }
                const int h = N / 2;
                s_w[pair * PS + (h >> S::L2) * (N2 + 1) + (h & (N2 - 1))] = make_float2(a, a);
            }
        }
        if (nyq_lane) {
            int nshort = 0;
            while (nshort < nring && c_half[nshort] < N / 2) ++nshort;
            for (int pair = nl; pair < NP; pair += 32)
#pragma unroll
                for (int cl = 0; cl < 8; ++cl) s_nyq[pair][cl] = 0.f;
            __syncwarp();
            // lane = (pair, ring parity): products of its rings go to the ring's length class
            for (int item = nl; item < 2 * NP; item += 32) {
                const int pair = item % NP, par = item / NP, m = pair / TN, n = pair % TN;
                const float2* dm = d2 + (size_t)(m >> 2) * gstride * 2;
                for (int base = par; base < nshort; base += 16) {
                    float cvv[8], dvv[8]; int cls[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = base + 2 * u;
                        const bool ok = i < nshort;
                        const int h = c_half[ok ? i : 0];
                        cls[u] = (29 - __clz(h)) & 7;                       // log2(half) - 2
                        cvv[u] = ok ? __ldg(c2 + cra_spec_idx(c_coff[ok ? i : 0], h, n, h)).x : 0.f;
                        dvv[u] = ok ? __ldg(dm + cra_spec_idx(c_coff[ok ? i : 0], h, m, h)).x : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (base + 2 * u < nshort) atomicAdd(&s_nyq[pair][cls[u]], cvv[u] * dvv[u]);
                }
            }
        }
    }
    __syncthreads();
    if (nyq_lane) {
        for (int pair = nl; pair < NP; pair += 32) {
            float2* w = s_w + pair * PS;
            int cur = -1;
            for (int i = 0; i < nring && c_half[i] < N / 2; ++i) {
                const int h = c_half[i];
                if (h == cur) continue;
                cur = h;
                const float a = s_nyq[pair][(29 - __clz(h)) & 7];
                const int hh = N - h;
                float2* p0 = w + (h >> S::L2) * (N2 + 1) + (h & (N2 - 1));
                float2* p1 = w + (hh >> S::L2) * (N2 + 1) + (hh & (N2 - 1));
                float2 v = *p0; v.x += a; v.y += a; *p0 = v;
                v = *p1; v.x += a; v.y += a; *p1 = v;
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
        fft_reg<N1, 1>(x);
#pragma unroll
        for (int j = 0; j < N1; ++j) {
            if (j == 0) { w[0] = x[0]; continue; }
            const float2 t = s_tw[j * N2 + n2];
            w[j * (N2 + 1)] = crafft::cmul(x[j], t);
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
        fft_reg<N2, 1>(x);
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
            if (live && row < nrows && ref < R) {
                if (qn >= qm) { cd.v = qn; cd.code = ref * 8192 + (mq + 1); }
                else          { cd.v = qm; cd.code = ref * 8192 + 4096 + (mt + 1); }
            } else { cd.v = -INFINITY; cd.code = -1; }
            s_pair[pair] = cd;
        }
    }
    __syncthreads();
    if (tid < TM && live) {
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

// ---- scalar helpers on the device spectrum layout (finalize / test entry) -------------------
__device__ __forceinline__ float2 spec_at(const float2* __restrict__ base, int nc, int row, int coff, int half, int k)
{
    return base[(size_t)(row >> 2) * nc * 4 + cra_spec_idx(coff, half, row, k)];
}

// q_k and t_k of one (row, ref) pair at frequency k, 0 <= k <= N/2
__device__ void pair_freq(const float2* __restrict__ spec, int row, const float2* __restrict__ refspec, int ref,
                          const CraRingTab* __restrict__ tab, int k, float& zq_r, float& zq_i, float& zt_r, float& zt_i)
{
    float A = 0.f, B = 0.f, C = 0.f, D = 0.f;
    for (int i = tab->nring - 1; i >= 0 && k <= (tab->len[i] >> 1); --i) {
        const int co = tab->coff[i], hf = tab->len[i] >> 1;
        const float2 c = spec_at(refspec, tab->nc, ref, co, hf, k), d = spec_at(spec, tab->nc, row, co, hf, k);
        A = fmaf(c.x, d.x, A); B = fmaf(c.y, d.y, B); C = fmaf(c.x, d.y, C); D = fmaf(c.y, d.x, D);
    }
    zq_r = A + B; zq_i = D - C; zt_r = A - B; zt_i = -C - D;
}

// ---- the same on the fragment layout (CRA_FMT_FRAG): value = bf16 hi + bf16 lo -------------------
__device__ __forceinline__ float frag_val(const uint4& u, int j)
{
    const unsigned h = (j < 2) ? u.x : u.y, l = (j < 2) ? u.z : u.w;
    const unsigned sh = (j & 1) ? 0u : 16u;
    return __uint_as_float((h << sh) & 0xffff0000u) + __uint_as_float((l << sh) & 0xffff0000u);
}

__device__ void pair_freq_frag(const unsigned char* __restrict__ rowp, const unsigned char* __restrict__ refp,
                               const CraFragTab& frag, int k, float& zq_r, float& zq_i, float& zt_r, float& zt_i)
{
    float A = 0.f, B = 0.f, C = 0.f, D = 0.f;
    const int c1 = frag.koff[k + 1];
    for (int gc = frag.koff[k]; gc < c1; ++gc)
        for (int t = 0; t < 4; ++t) {
            const uint4* dp = reinterpret_cast<const uint4*>(rowp + (size_t)gc * 128 + t * 32);
            const uint4* cp = reinterpret_cast<const uint4*>(refp + (size_t)gc * 128 + t * 32);
            // row: A-operand order hi{re01, im01, re23, im23}, lo{same}; reference: [re unit | im unit]
            const uint4 dhi = dp[0], dlo = dp[1], cre = cp[0], cim = cp[1];
            const uint4 dre = frag.unit_rows ? dhi : make_uint4(dhi.x, dhi.z, dlo.x, dlo.z);
            const uint4 dim = frag.unit_rows ? dlo : make_uint4(dhi.y, dhi.w, dlo.y, dlo.w);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float dx = frag_val(dre, j), dy = frag_val(dim, j), cx = frag_val(cre, j), cy = frag_val(cim, j);
                A = fmaf(cx, dx, A); B = fmaf(cy, dy, B); C = fmaf(cx, dy, C); D = fmaf(cy, dx, D);
            }
        }
    zq_r = A + B; zq_i = D - C; zt_r = A - B; zt_i = -C - D;
}

// A, B, C, D ring sums of one (row, ref) pair at frequency k in DOUBLE precision (the operands are exact
// in FP32, so every product is exact in double): q_k = (A+B) + i(D-C), t_k = (A-B) - i(C+D)
__device__ void pair_freq_d(const float2* __restrict__ spec, int row, const float2* __restrict__ refspec, int ref,
                            const CraRingTab* __restrict__ tab, int k, double& zq_r, double& zq_i, double& zt_r, double& zt_i)
{
    double A = 0., B = 0., C = 0., D = 0.;
    for (int i = tab->nring - 1; i >= 0 && k <= (tab->len[i] >> 1); --i) {
        const int co = tab->coff[i], hf = tab->len[i] >> 1;
        const float2 c = spec_at(refspec, tab->nc, ref, co, hf, k), d = spec_at(spec, tab->nc, row, co, hf, k);
        A += (double)c.x * d.x; B += (double)c.y * d.y; C += (double)c.x * d.y; D += (double)c.y * d.x;
    }
    zq_r = A + B; zq_i = D - C; zt_r = A - B; zt_i = -C - D;
}
__device__ void pair_freq_frag_d(const unsigned char* __restrict__ rowp, const unsigned char* __restrict__ refp,
                                 const CraFragTab& frag, int k, double& zq_r, double& zq_i, double& zt_r, double& zt_i)
{
    double A = 0., B = 0., C = 0., D = 0.;
    const int c1 = frag.koff[k + 1];
    for (int gc = frag.koff[k]; gc < c1; ++gc)
        for (int t = 0; t < 4; ++t) {
            const uint4* dp = reinterpret_cast<const uint4*>(rowp + (size_t)gc * 128 + t * 32);
            const uint4* cp = reinterpret_cast<const uint4*>(refp + (size_t)gc * 128 + t * 32);
            const uint4 dhi = dp[0], dlo = dp[1], cre = cp[0], cim = cp[1];
            const uint4 dre = frag.unit_rows ? dhi : make_uint4(dhi.x, dhi.z, dlo.x, dlo.z);
            const uint4 dim = frag.unit_rows ? dlo : make_uint4(dhi.y, dhi.w, dlo.y, dlo.w);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double dx = frag_val(dre, j), dy = frag_val(dim, j), cx = frag_val(cre, j), cy = frag_val(cim, j);
                A += cx * dx; B += cy * dy; C += cx * dy; D += cy * dx;
            }
        }
    zq_r = A + B; zq_i = D - C; zt_r = A - B; zt_i = -C - D;
}

// One warp per particle: the winner over (row, reference) in the reference's visit order from the candidates of
// the CCF kernel, then BOTH correlation curves of that pair in double precision straight from the spectra -- the
// arithmetic of EMAN2's Crosrng_ms (double accumulation, fftr_d): ">=" argmax over the lags (the last maximum
// wins), "qn >= qm" for the mirror, prb1d on the 7 samples around the maximum, ang_n, and the final rotation of
// the shift.  norm / tref: deferred Normalize_ring of the fragment kernels (null: the spectra are normalised).
template <int FMT>
__global__ void __launch_bounds__(128)
finalize_kernel(const float2* __restrict__ spec, const float2* __restrict__ refspec, int R,
                const CraRingTab* __restrict__ tab, const CraCand* __restrict__ cand, int ntile_n,
                CraRowMap map, CraResult* __restrict__ out, CraFragTab frag, const double2* __restrict__ twd,
                const float2* __restrict__ norm, const float* __restrict__ tref)
{
    extern __shared__ __align__(16) double2 s_xd[];           // per warp: N points, W = q + i t, then the two curves
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int p = blockIdx.x * (blockDim.x >> 5) + wib;
    if (p >= map.np) return;
    const int N = tab->maxrin, LN = 31 - __clz(N);
    double2* x = s_xd + (size_t)wib * N;
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

    for (int k = lane; k <= N / 2; k += 32) {
        double4 v;
        if (FMT == CRA_FMT_FRAG) {
            const size_t rb = cra_frag_row_bytes(frag.nch);
            pair_freq_frag_d(reinterpret_cast<const unsigned char*>(spec) + (size_t)row * rb,
                             reinterpret_cast<const unsigned char*>(refspec) + (size_t)iref * rb, frag, k, v.x, v.y, v.z, v.w);
        } else pair_freq_d(spec, row, refspec, iref, tab, k, v.x, v.y, v.z, v.w);
        // Hermitian extension of W = Q + i T (q and t are real sequences), stored bit-reversed for the in-place FFT:
        // W[k] = (Qr - Ti, Qi + Tr), W[N-k] = conj(Q) + i conj(T) = (Qr + Ti, Tr - Qi)
        x[__brev((unsigned)k) >> (32 - LN)] = make_double2(v.x - v.w, v.y + v.z);
        if (k != 0 && k != N / 2) x[__brev((unsigned)(N - k)) >> (32 - LN)] = make_double2(v.x + v.w, v.z - v.y);
    }
    __syncwarp();
    // One complex inverse FFT of length N in double (radix 2, decimation in time, in place) gives both curves at
    // every lag: x[m] = (q[m], t[m]) * N -- EMAN2's fftr_d works in double as well; this replaces the direct
    // N x (N/2 + 1) sums (8.2 -> ~2 ms per 100k particles).
    for (int s = 0; s < LN; ++s) {
        const int half = 1 << s;
        for (int j = lane; j < N / 2; j += 32) {
            const int pos = j & (half - 1), i0 = ((j >> s) << (s + 1)) + pos, i1 = i0 + half;
            const double2 w = twd[pos << (LN - 1 - s)];          // exp(+2 pi i pos / (2 half))
            const double2 a = x[i0], b = x[i1];
            const double2 bw = make_double2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
            x[i0] = make_double2(a.x + bw.x, a.y + bw.y);
            x[i1] = make_double2(a.x - bw.x, a.y - bw.y);
        }
        __syncwarp();
    }
    const double invN = 1.0 / (double)N;
    // scan order j = 1..maxrin with ">=": the later lag wins a tie
    double bq = -INFINITY, bt = -INFINITY; int mq = -1, mt = -1;
    for (int m = lane; m < N; m += 32) {
        const double2 v = make_double2(x[m].x * invN, x[m].y * invN);
        if (v.x >= bq) { bq = v.x; mq = m; }
        if (v.y >= bt) { bt = v.y; mt = m; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double oq = __shfl_xor_sync(0xffffffffu, bq, o); const int omq = __shfl_xor_sync(0xffffffffu, mq, o);
        const double ot = __shfl_xor_sync(0xffffffffu, bt, o); const int omt = __shfl_xor_sync(0xffffffffu, mt, o);
        if (oq > bq || (oq == bq && omq > mq)) { bq = oq; mq = omq; }
        if (ot > bt || (ot == bt && omt > mt)) { bt = ot; mt = omt; }
    }
    const int mirror = (bq >= bt) ? 0 : 1;
    const int jtot = (mirror ? mt : mq) + 1;                    // 1-based lag of the maximum
    double t7 = 0.;
    if (lane < 7) {
        const double2 v = x[(jtot - 1 + lane - 3 + N) & (N - 1)];
        t7 = (mirror ? v.y : v.x) * invN;
    }
    double b[7];
#pragma unroll
    for (int s = 0; s < 7; ++s) b[s] = __shfl_sync(0xffffffffu, t7, s);
    if (lane == 0) {
        const double c2 = 49. * b[0] + 6. * b[1] - 21. * b[2] - 32. * b[3] - 27. * b[4] - 6. * b[5] + 31. * b[6];
        const double c3 = 5. * b[0] - 3. * b[2] - 4. * b[3] - 3. * b[4] + 5. * b[6];
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
        double peak = mirror ? bt : bq;
        if (norm) {           // deferred Normalize_ring: every lag moves by -avg * tref / N, then 1/sigma
            const float2 nm = norm[row];
            peak = (peak - (double)nm.x * (double)tref[iref] / N) * (double)nm.y;
        }
        res.ang = ang; res.sxs = sx * co - sy * so; res.sys = sx * so + sy * co;
        res.mirror = mirror; res.iref = iref; res.peak = (float)peak; res.sx = sx; res.sy = sy;
        out[p] = res;
    }
}

// Test entry: full q/t curves of one pair by direct evaluation.
template <int FMT>
__global__ void ccf_curves_kernel(const float2* __restrict__ spec, int row, const float2* __restrict__ refspec, int ref,
                                  const CraRingTab* __restrict__ tab, float* __restrict__ q, float* __restrict__ t, CraFragTab frag)
{
    extern __shared__ float4 s_z[];   // N/2+1 entries: (qr, qi, tr, ti)
    const int N = tab->maxrin;
    for (int k = threadIdx.x; k <= N / 2; k += blockDim.x) {
        float4 z;
        if (FMT == CRA_FMT_FRAG) {
            const size_t rb = cra_frag_row_bytes(frag.nch);
            pair_freq_frag(reinterpret_cast<const unsigned char*>(spec) + (size_t)row * rb,
                           reinterpret_cast<const unsigned char*>(refspec) + (size_t)ref * rb, frag, k, z.x, z.y, z.z, z.w);
        } else pair_freq(spec, row, refspec, ref, tab, k, z.x, z.y, z.z, z.w);
        s_z[k] = z;
    }
    __syncthreads();
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
        double aq = 0, at = 0;
        for (int k = 0; k <= N / 2; ++k) {
            const int ph = (int)(((long long)k * m) % N);
            double sn, cs; sincospi(2.0 * (double)ph / (double)N, &sn, &cs);
            const float4 z = s_z[k];
            const double wgt = (k == 0 || k == N / 2) ? 1.0 : 2.0;
            aq += wgt * (z.x * cs - z.y * sn);
            at += wgt * (z.z * cs - z.w * sn);
        }
        q[m] = (float)(aq / N);
        t[m] = (float)(at / N);
    }
}

// upload the ring table to constant memory when it differs from the resident one
int bind_ring_table(const CraRingTab& h, cudaStream_t st)
{
    static int cur_coff[CRA_MAX_RINGS], cur_half[CRA_MAX_RINGS], cur_n = -1, cur_dev = -1;
    int half[CRA_MAX_RINGS];
    for (int i = 0; i < h.nring; ++i) half[i] = h.len[i] >> 1;
    int dev = 0; cudaGetDevice(&dev);
    if (cur_n == h.nring && cur_dev == dev && memcmp(cur_coff, h.coff, sizeof(int) * h.nring) == 0 &&
        memcmp(cur_half, half, sizeof(int) * h.nring) == 0) return 0;
    CRA_CUDA(cudaStreamSynchronize(st));
    CRA_CUDA(cudaMemcpyToSymbol(c_coff, h.coff, sizeof(int) * h.nring));
    CRA_CUDA(cudaMemcpyToSymbol(c_half, half, sizeof(int) * h.nring));
    // the copies come from pageable memory on the legacy stream and the kernel runs on a non-blocking stream:
    // make sure they have landed (diagnostic path; one-off per geometry and device)
    CRA_CUDA(cudaDeviceSynchronize());
    memcpy(cur_coff, h.coff, sizeof(int) * h.nring); memcpy(cur_half, half, sizeof(int) * h.nring);
    cur_n = h.nring; cur_dev = dev;
    return 0;
}

template <int LOG2N>
int launch_ccf_t(const float* spec, int nrows, const float* refspec, int R, const CraRingTab& h,
                 const float2* twid, CraCand* cand, int ntile_n, cudaStream_t st)
{
    using S = Shape<LOG2N>;
    constexpr int SUBM = S::SUBM, SUBN = S::SUBN, NSUB = S::NSUB;
    const size_t smem = ((size_t)NSUB * NP * S::PS + S::N) * sizeof(float2);
    if (cra_ensure_dyn_smem(reinterpret_cast<const void*>(&ccf_peak_kernel<LOG2N>), smem)) return 1;
    const long ntile_m = (nrows + TM - 1) / TM;
    const long ncta_m = (ntile_m + SUBM - 1) / SUBM, ncta_n = (ntile_n + SUBN - 1) / SUBN;
    const long nblk = ncta_m * ncta_n;
    if (nblk <= 0) return 0;
    if (nblk > 2147483647L) { cra_set_error("ccf grid too large; lower row_batch"); return 1; }
    ccf_peak_kernel<LOG2N><<<(unsigned)nblk, S::NT * NSUB, smem, st>>>(reinterpret_cast<const float4*>(spec), nrows,
                                                                       reinterpret_cast<const float4*>(refspec), R,
                                                                       h.nring, h.nc, twid, cand, ntile_n,
                                                                       (int)ntile_m, (int)ncta_n);
    CRA_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int cra_ccf_tile_n() { return TN; }

// twiddle table layout the CCF kernel expects: tw[j*N2 + n2] = exp(+2 pi i n2 j / N)
void cra_ccf_twiddles(int log2n, std::vector<float2>& tw)
{
    const int N = 1 << log2n, L1 = log2n / 2, N1 = 1 << L1, N2 = N / N1;
    tw.resize(N);
    for (int j = 0; j < N1; ++j)
        for (int n2 = 0; n2 < N2; ++n2) {
            const double a = 2.0 * M_PI * (double)((n2 * j) % N) / N;
            tw[j * N2 + n2] = make_float2((float)cos(a), (float)sin(a));
        }
}

int cra_launch_ccf(const float* spec, int nrows, const float* refspec, int R, const CraRingTab* tab,
                   const CraRingTab& htab, const float2* twid, CraCand* cand, int ntile_n, cudaStream_t st)
{
    (void)tab;
    if (bind_ring_table(htab, st)) return 1;
    switch (htab.log2n) {
        case 5:  return launch_ccf_t<5>(spec, nrows, refspec, R, htab, twid, cand, ntile_n, st);
        case 6:  return launch_ccf_t<6>(spec, nrows, refspec, R, htab, twid, cand, ntile_n, st);
        case 7:  return launch_ccf_t<7>(spec, nrows, refspec, R, htab, twid, cand, ntile_n, st);
        case 8:  return launch_ccf_t<8>(spec, nrows, refspec, R, htab, twid, cand, ntile_n, st);
        case 9:  return launch_ccf_t<9>(spec, nrows, refspec, R, htab, twid, cand, ntile_n, st);
        case 10: return launch_ccf_t<10>(spec, nrows, refspec, R, htab, twid, cand, ntile_n, st);
        default: cra_set_error("maxrin must be a power of two in [32, 1024]"); return 1;
    }
}

int cra_launch_finalize(const float* spec, const float* refspec, int R, const CraRingTab* tab, const CraRingTab& htab,
                        const CraCand* cand, int ntile_n, CraRowMap map, CraResult* out, int fmt, const CraFragTab& frag,
                        const double2* twd, const float2* norm, const float* tref, cudaStream_t st)
{
    if (map.np <= 0) return 0;
    const int wpb = 4;
    const size_t smem = (size_t)wpb * htab.maxrin * sizeof(double2);
    const float2* s2 = reinterpret_cast<const float2*>(spec); const float2* r2 = reinterpret_cast<const float2*>(refspec);
    if (smem > 48 * 1024 && cra_ensure_dyn_smem(fmt == CRA_FMT_FRAG ? reinterpret_cast<const void*>(&finalize_kernel<CRA_FMT_FRAG>)
                                                                     : reinterpret_cast<const void*>(&finalize_kernel<CRA_FMT_F32>), smem)) return 1;
    if (fmt == CRA_FMT_FRAG)
        finalize_kernel<CRA_FMT_FRAG><<<(map.np + wpb - 1) / wpb, wpb * 32, smem, st>>>(s2, r2, R, tab, cand, ntile_n, map, out, frag, twd, norm, tref);
    else
        finalize_kernel<CRA_FMT_F32><<<(map.np + wpb - 1) / wpb, wpb * 32, smem, st>>>(s2, r2, R, tab, cand, ntile_n, map, out, frag, twd, nullptr, nullptr);
    CRA_CUDA(cudaGetLastError());
    return 0;
}

int cra_launch_ccf_curves(const float* spec, int row, const float* refspec, int ref, const CraRingTab* tab,
                          const CraRingTab& htab, float* q, float* t, int fmt, const CraFragTab& frag, cudaStream_t st)
{
    const float2* s2 = reinterpret_cast<const float2*>(spec); const float2* r2 = reinterpret_cast<const float2*>(refspec);
    const size_t sm = (htab.maxrin / 2 + 1) * sizeof(float4);
    if (fmt == CRA_FMT_FRAG) ccf_curves_kernel<CRA_FMT_FRAG><<<1, 256, sm, st>>>(s2, row, r2, ref, tab, q, t, frag);
    else ccf_curves_kernel<CRA_FMT_F32><<<1, 256, sm, st>>>(s2, row, r2, ref, tab, q, t, frag);
    CRA_CUDA(cudaGetLastError());
    return 0;
}
