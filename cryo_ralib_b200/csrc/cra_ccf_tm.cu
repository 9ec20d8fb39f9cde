// cra_ccf_tm.cu -- Crosrng_ms ring contraction on the tensor cores with the correlation spectrum
// staged in TENSOR MEMORY, fused with the inverse FFT and the peak search (EMAN2 Util::Crosrng_ms +
// the best-of loop of Util::multiref_polar_ali_2d; reference call site test_mref.py:200-201;
// replaces cu_ccf_mult_m + cuFFT C2R + cu_max_idx_batch, cuda/gpu_aln_noref.cu:1009-1143,
// :2198-2206, :1305-1346).
//
// Same arithmetic as cra_ccf_mma.cu (mma.sync.m16n8k16 split-bf16 x3, W = q + i t, one complex
// inverse FFT per pair, ">=" argmax).  What changes is where W lives.  The 2 KB of W per
// (row, reference) pair cap a shared-memory tile at ~100 pairs and ONE resident CTA per SM, so the
// L2-latency-bound contraction and the FP32-bound inverse FFT can never overlap.  Here a warp writes
// its finished frequencies to TMEM with tcgen05.st: in the mma accumulator layout lane (g, t) owns
// the pairs (row g, reference 4j + t) in EVERY warp, and a thread may only touch the TMEM lanes of its
// warp's quadrant (warp % 4), so the frequencies are dealt to the quadrants by residue
// n2 = k mod N2 (and N2 - n2: the Hermitian partner W[N-k] of the same MMA result).  Pass 1 of the
// inverse FFT (N1-point DFTs over k = n1 N2 + n2 at fixed n2) then reads nothing but the thread's own
// TMEM lane (tcgen05.ld .32x32b), and only its output goes to shared memory, one 4-reference
// batch (32 pairs, 70 KB) at a time, for pass 2 and the argmax.  A CTA is 8 warps, 8 rows x 8
// references, 256 TMEM columns and 74 KB of shared memory: TWO CTAs are resident per SM and the
// contraction of one runs under the inverse FFT of the other.
#include "cra_common.cuh"
#include "cra_fft.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <algorithm>

namespace {

#ifndef CRA_TM_WARPS
#define CRA_TM_WARPS 8
#endif
constexpr int kWarps = CRA_TM_WARPS;      // a multiple of 4: kWarps / 4 warps share each TMEM quadrant
constexpr int kThreads = kWarps * 32;
constexpr int kMaxItems = 640;          // chunk items per warp
constexpr int kMaxFreq = 160;           // frequencies per warp
constexpr int kPfDist = 40;             // row blocks between an L2 prefetch and its use

// per-warp work lists (global, staged in shared memory by every CTA):
//   items: gc | last << 23   (gc = chunk index within a row)
//   flush: TMEM column (relative to the pair slot) of W[k] | column of W[N-k] << 12 | has_partner << 24
__device__ int g_items[kWarps][kMaxItems];
__device__ int g_nitems[kWarps];
__device__ int g_flush[kWarps][kMaxFreq];
__device__ int g_res[4][8];             // residues n2 of quadrant q, in TMEM order (N2 / 4 of them)

using crafft::fft_reg;

template <int LOG2N>
struct TShape {
    static constexpr int N = 1 << LOG2N;
    static constexpr int L1 = LOG2N / 2;
    static constexpr int L2 = LOG2N - L1;
    static constexpr int N1 = 1 << L1;
    static constexpr int N2 = 1 << L2;
    static constexpr int PS = N1 * (N2 + 1) + ((N1 * (N2 + 1)) % 2 == 0 ? 1 : 0);   // odd float2 stride of one pair
    static constexpr int NJ = 2;                        // reference quads per CTA (maxrin 512: all 512 columns, one CTA per SM)
    static constexpr int RQ = N2 / 4;                   // residues per quadrant
    static constexpr int JCOLS = N / 2;                 // TMEM columns of one pair slot in one quadrant = RQ * N1 * 2
    static constexpr int COLS = (NJ * JCOLS < 32) ? 32 : NJ * JCOLS;
};

__device__ __forceinline__ void mma_bf16(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct Frag8 { unsigned w[8]; };
__device__ __forceinline__ Frag8 ldg256(const unsigned char* p)
{
    Frag8 f;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(f.w[0]), "=r"(f.w[1]), "=r"(f.w[2]), "=r"(f.w[3]), "=r"(f.w[4]), "=r"(f.w[5]), "=r"(f.w[6]), "=r"(f.w[7])
                 : "l"(p));
    return f;
}

__device__ __forceinline__ bool better(float v, int m, float bv, int bm)
{   // ">=" scan order semantics: larger value wins, ties go to the later index
    return (v > bv) || (v == bv && m > bm);
}

// ---- tensor memory -------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st2(unsigned taddr, float a, float b)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)) : "memory");
}
// NC consecutive columns of this thread's own TMEM lane -> NC/2 complex values
template <int NC> __device__ __forceinline__ void tmem_ld(unsigned taddr, float2 (&x)[NC / 2]);
template <> __device__ __forceinline__ void tmem_ld<8>(unsigned taddr, float2 (&x)[4])
{
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
}
template <> __device__ __forceinline__ void tmem_ld<16>(unsigned taddr, float2 (&x)[8])
{
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
}
template <> __device__ __forceinline__ void tmem_ld<32>(unsigned taddr, float2 (&x)[16])
{
    unsigned r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
}

template <int NJ>
struct Operands { Frag8 a; uint4 b[NJ]; };

template <int LOG2N>
__global__ void __launch_bounds__(kThreads, 2)
ccf_tm_kernel(const unsigned char* __restrict__ spec, int nrows, const unsigned char* __restrict__ refspec, int R,
              size_t row_bytes, const float2* __restrict__ twid, CraCand* __restrict__ cand,
              int nquad, int ncta_n, const float2* __restrict__ norm, const float* __restrict__ tref, int istride, int fstride,
              long ntiles)
{
    using S = TShape<LOG2N>;
    constexpr int N = S::N, N1 = S::N1, N2 = S::N2, PS = S::PS, NJ = S::NJ, RQ = S::RQ, JCOLS = S::JCOLS;
    extern __shared__ __align__(16) float2 s_dyn[];
    float2* s_y = s_dyn;                      // 32 pairs * PS : pass-1 output of one reference quad
    float2* s_tw = s_dyn + 32 * PS;           // N : s_tw[j*N2 + n2] = exp(+2 pi i n2 j / N)
    int* s_items = reinterpret_cast<int*>(s_tw + N);            // kWarps * istride chunk items
    int* s_flush = s_items + kWarps * istride;                  // kWarps * fstride completed frequencies
    __shared__ CraCand s_pair[32 * NJ];
    __shared__ unsigned s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int quad = warp & 3;                                  // TMEM lane quadrant of this warp
    const int qbase = nquad / ncta_n, qrem = nquad % ncta_n;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((unsigned)__cvta_generic_to_shared(&s_tmem)), "n"(S::COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < N; i += kThreads) s_tw[i] = twid[i];
    {
        const int nit = g_nitems[warp];
        for (int i = lane; i < nit; i += 32) s_items[warp * istride + i] = g_items[warp][i];
        for (int i = lane; i < fstride; i += 32) s_flush[warp * fstride + i] = g_flush[warp][i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // this thread's TMEM lane: bits 31:16 lane, 15:0 column
    const unsigned tbase = s_tmem + ((unsigned)(quad * 32) << 16);

    // Persistent CTA: the work lists, the twiddles and the TMEM allocation above are set up once, then
    // the CTA walks the tiles with the grid stride.  Reference tile fastest, so that the CTAs resident
    // at any moment share the row spectra in L2.
#pragma unroll 1
    for (int tile = blockIdx.x; tile < (int)ntiles; tile += gridDim.x) {
    {
    const int cm = tile / ncta_n, cn = tile - cm * ncta_n;
    const int nj = qbase + (cn < qrem ? 1 : 0);                 // reference quads of this tile (<= NJ)
    const int q0 = cn * qbase + min(cn, qrem);
    const int row0 = (int)(cm * 8);
    if (cn == 0) {                            // pull a future row block into L2 (see cra_ccf_mma.cu)
        const long r0 = (long)(cm + kPfDist) * 8;
        if (r0 < nrows) {
            const long r1 = min((long)nrows, r0 + 8);
            const unsigned char* p0 = spec + (size_t)r0 * row_bytes;
            const size_t nline = (size_t)(r1 - r0) * row_bytes / 128;
            for (size_t i = tid; i < nline; i += kThreads)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(p0 + i * 128));
        }
    }
    // the previous tile's tcgen05.ld (pass 1) were ordered before its barriers
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- contraction: this warp's frequencies, chunk by chunk ---------------------------------
    {
        int rrow = row0 + g;
        if (rrow >= nrows) rrow = nrows - 1;
        const unsigned char* pa = spec + (size_t)rrow * row_bytes + t * 32;
        const unsigned char* pb[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int jj = (j < nj) ? j : 0;
            pb[j] = refspec + (size_t)(4 * (q0 + jj) + (g >> 1)) * row_bytes + t * 32 + (g & 1) * 16;
        }
        float acc[NJ][4];
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;

        const int nit = g_nitems[warp];
        const int* items = s_items + warp * istride;
        const int* fl = s_flush + warp * fstride;

#define CRA_LOAD_OPS(O, item)                                                            \
        { const size_t off_ = (size_t)((item) & 8191) * 128;                             \
          O.a = ldg256(pa + off_);                                                       \
          _Pragma("unroll") for (int j_ = 0; j_ < NJ; ++j_)                              \
              O.b[j_] = __ldg(reinterpret_cast<const uint4*>(pb[j_] + off_)); }
#define CRA_COMPUTE(O, item)                                                             \
        { _Pragma("unroll") for (int j_ = 0; j_ < NJ; ++j_)                              \
              mma_bf16(acc[j_], O.a.w[0], O.a.w[1], O.a.w[2], O.a.w[3], O.b[j_].z, O.b[j_].w);   /* a_hi b_lo */ \
          _Pragma("unroll") for (int j_ = 0; j_ < NJ; ++j_)                              \
              mma_bf16(acc[j_], O.a.w[4], O.a.w[5], O.a.w[6], O.a.w[7], O.b[j_].x, O.b[j_].y);   /* a_lo b_hi */ \
          _Pragma("unroll") for (int j_ = 0; j_ < NJ; ++j_)                              \
              mma_bf16(acc[j_], O.a.w[0], O.a.w[1], O.a.w[2], O.a.w[3], O.b[j_].x, O.b[j_].y);   /* a_hi b_hi */ \
          if ((item) >> 23) flush_freq(); }

        auto flush_freq = [&]() {
            const int f = *fl++;
            const unsigned c0 = f & 4095, c1 = (f >> 12) & 4095;
            const bool two = (f >> 24) != 0;                    // warp-uniform
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                // c0 = A (re.re), c1 = D (row re . ref im), c2 = C (row im . ref re), c3 = B (im.im)
                const float A = acc[j][0], D = acc[j][1], C = acc[j][2], B = acc[j][3];
                // s = (A+B, A-B), tv = (C+D, D-C);  W[k] = s + tv,  W[N-k] = s - tv
                const float sx = A + B, sy = A - B, tx = C + D, ty = D - C;
                tmem_st2(tbase + j * JCOLS + c0, sx + tx, sy + ty);
                if (two) tmem_st2(tbase + j * JCOLS + c1, sx - tx, sy - ty);
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
            }
        };

#if CRA_TM_WARPS <= 8
        // Four operand sets, loaded two chunks at a time.  ptxas keeps every operand LDG on one hardware
        // scoreboard and drains it before the oldest set is used (profiles/README.md), so what a load gets
        // to hide behind is the work issued between its batch and the next drain: issuing the loads in
        // pairs (drain, load chunks i+2 and i+3, multiply chunks i and i+1) doubles that window.
        Operands<NJ> o0, o1, o2, o3;
        if (nit > 0) CRA_LOAD_OPS(o0, items[0]);
        if (nit > 1) CRA_LOAD_OPS(o1, items[1]);
        for (int i = 0; i < nit; i += 4) {
            const int e0 = items[i];
            const int e1 = (i + 1 < nit) ? items[i + 1] : 0;
            if (i + 2 < nit) CRA_LOAD_OPS(o2, items[i + 2]);
            if (i + 3 < nit) CRA_LOAD_OPS(o3, items[i + 3]);
            CRA_COMPUTE(o0, e0);
            if (i + 1 >= nit) break;
            CRA_COMPUTE(o1, e1);
            if (i + 2 >= nit) break;
            const int e2 = items[i + 2];
            const int e3 = (i + 3 < nit) ? items[i + 3] : 0;
            if (i + 4 < nit) CRA_LOAD_OPS(o0, items[i + 4]);
            if (i + 5 < nit) CRA_LOAD_OPS(o1, items[i + 5]);
            CRA_COMPUTE(o2, e2);
            if (i + 3 >= nit) break;
            CRA_COMPUTE(o3, e3);
        }
#else
        // three operand sets (register budget of the wider CTA)
        Operands<NJ> o0, o1, o2;
        if (nit > 0) CRA_LOAD_OPS(o0, items[0]);
        if (nit > 1) CRA_LOAD_OPS(o1, items[1]);
        for (int i = 0; i < nit; i += 3) {
            const int e0 = items[i];
            if (i + 2 < nit) CRA_LOAD_OPS(o2, items[i + 2]);
            CRA_COMPUTE(o0, e0);
            if (i + 1 >= nit) break;
            const int e1 = items[i + 1];
            if (i + 3 < nit) CRA_LOAD_OPS(o0, items[i + 3]);
            CRA_COMPUTE(o1, e1);
            if (i + 2 >= nit) break;
            const int e2 = items[i + 2];
            if (i + 4 < nit) CRA_LOAD_OPS(o1, items[i + 4]);
            CRA_COMPUTE(o2, e2);
        }
#endif
#undef CRA_LOAD_OPS
#undef CRA_COMPUTE
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    {
    // tile scalars re-derived here instead of kept in registers across the contraction
    int tile2 = tile; asm volatile("" : "+r"(tile2));
    const int cm = tile2 / ncta_n, cn = tile2 - cm * ncta_n;
    const int nj = qbase + (cn < qrem ? 1 : 0);
    const int q0 = cn * qbase + min(cn, qrem);
    const int row0 = cm * 8;

    // ---- inverse FFT, one reference quad (32 pairs: lane = pair) at a time -----------------------
    for (int j = 0; j < nj; ++j) {
        // pass 1: (pair = lane, residue n2): N1-point DFT over n1 straight from this lane's TMEM, twiddle
        for (int ri = warp >> 2; ri < RQ; ri += kWarps / 4) {
            const int n2 = g_res[quad][ri];
            float2 x[N1];
            tmem_ld<2 * N1>(tbase + j * JCOLS + ri * (2 * N1), x);
            fft_reg<N1, 1>(x);
            float2* w = s_y + lane * PS + n2;
#pragma unroll
            for (int k1 = 0; k1 < N1; ++k1) {
                if (k1 == 0) { w[0] = x[0]; continue; }
                const float2 tw = s_tw[k1 * N2 + n2];
                w[k1 * (N2 + 1)] = crafft::cmul(x[k1], tw);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // pass 2: (pair, k1): N2-point DFT over n2 -> X[k1 + N1*k2]; argmax over lags
        for (int item = tid; item < 32 * N1; item += kThreads) {
            const int pi = item / N1, k1 = item - pi * N1;
            const float2* w = s_y + pi * PS + k1 * (N2 + 1);
            float2 x[N2];
#pragma unroll
            for (int jj = 0; jj < N2; ++jj) x[jj] = w[jj];
            fft_reg<N2, 1>(x);
            float bq = -INFINITY, bt = -INFINITY; int mq = -1, mt = -1;
#pragma unroll
            for (int jj = 0; jj < N2; ++jj) {
                const int m = k1 + N1 * jj;
                if (x[jj].x >= bq) { bq = x[jj].x; mq = m; }
                if (x[jj].y >= bt) { bt = x[jj].y; mt = m; }
            }
            // the N1 lanes of one pair are consecutive and aligned inside a warp (N1 <= 32)
#pragma unroll
            for (int o = N1 >> 1; o > 0; o >>= 1) {
                float oq = __shfl_xor_sync(0xffffffffu, bq, o); int omq = __shfl_xor_sync(0xffffffffu, mq, o);
                float ot = __shfl_xor_sync(0xffffffffu, bt, o); int omt = __shfl_xor_sync(0xffffffffu, mt, o);
                if (better(oq, omq, bq, mq)) { bq = oq; mq = omq; }
                if (better(ot, omt, bt, mt)) { bt = ot; mt = omt; }
            }
            if (k1 == 0) {
                const int row = row0 + (pi >> 2), ref = 4 * (q0 + j) + (pi & 3);     // pair = lane (g, t)
                CraCand cd;
                if (row < nrows && ref < R) {
                    // deferred Normalize_ring (cra_common.cuh): every lag moves by -avg * tref[ref]
                    const float2 nm = norm[row];
                    const float sc = nm.y / (float)N, dc = nm.x * tref[ref];
                    const float qn = (bq - dc) * sc, qm = (bt - dc) * sc;
                    if (qn >= qm) { cd.v = qn; cd.code = ref * 8192 + (mq + 1); }
                    else          { cd.v = qm; cd.code = ref * 8192 + 4096 + (mt + 1); }
                } else { cd.v = -INFINITY; cd.code = -1; }
                s_pair[j * 32 + pi] = cd;
            }
        }
        __syncthreads();
    }
    if (tid < 8) {
        const int row = row0 + tid;
        if (row < nrows) {
            CraCand best; best.v = -INFINITY; best.code = -1;
            for (int j = 0; j < nj; ++j)
                for (int c = 0; c < 4; ++c) {
                    const CraCand cd = s_pair[j * 32 + tid * 4 + c];
                    if (cd.code >= 0 && cd.v >= best.v) best = cd;
                }
            cand[(size_t)row * ncta_n + cn] = best;
        }
    }
    }
    }   // tile loop
    // every tcgen05.ld has completed (wait::ld) before the barriers above
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem), "n"(S::COLS) : "memory");
    }
}

struct Sched { int nring = -1, maxrin = -1, dev = -1, istride = 0, fstride = 0; std::vector<int> len; };
Sched g_sched;

// Deal the frequencies to the TMEM quadrants by residue n2 = k mod N2 (a residue and its negative
// together), balance each quadrant's frequencies over its two warps, and upload the lists.
int bind_schedule(const CraRingTab& h, const std::vector<int>& koff, cudaStream_t st)
{
    int dev = 0; cudaGetDevice(&dev);
    std::vector<int> len(h.len, h.len + h.nring);
    if (g_sched.nring == h.nring && g_sched.maxrin == h.maxrin && g_sched.dev == dev && g_sched.len == len) return 0;
    const int N = h.maxrin, nk = N / 2 + 1;
    const int L1 = h.log2n / 2, L2 = h.log2n - L1, N1 = 1 << L1, N2 = 1 << L2, RQ = N2 / 4;
    if (RQ < 1 || RQ > 8) { cra_set_error("tensor-memory CCF kernel: unsupported maxrin"); return 1; }
    auto nch = [&](int k) { return koff[k + 1] - koff[k]; };
    // residue classes {n2, N2 - n2} and their chunk loads
    std::vector<std::vector<int>> cls;
    for (int n2 = 0; n2 <= N2 / 2; ++n2) {
        std::vector<int> c{n2};
        if (n2 != 0 && n2 != N2 / 2) c.push_back(N2 - n2);
        cls.push_back(c);
    }
    auto cls_load = [&](const std::vector<int>& c) { int s = 0; for (int k = 0; k < nk; ++k) for (int r : c) if ((k & (N2 - 1)) == r) s += nch(k); return s; };
    // quadrant q must receive exactly RQ residues: the two single-residue classes (0 and N2/2) go together
    std::vector<std::vector<int>> qres(4);
    std::vector<int> qload(4, 0);
    std::vector<int> order;
    for (size_t i = 0; i < cls.size(); ++i) if (cls[i].size() == 2) order.push_back((int)i);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return cls_load(cls[a]) > cls_load(cls[b]); });
    qres[0] = {0, N2 / 2}; qload[0] = cls_load(cls[0]) + cls_load(cls[N2 / 2]);
    for (int ci : order) {
        int best = -1;
        for (int q = 0; q < 4; ++q) if ((int)qres[q].size() + 2 <= RQ && (best < 0 || qload[q] < qload[best])) best = q;
        if (best < 0) { cra_set_error("tensor-memory CCF kernel: residue assignment failed"); return 1; }
        qres[best].push_back(cls[ci][0]); qres[best].push_back(cls[ci][1]);
        qload[best] += cls_load(cls[ci]);
    }
    if (N2 == 4) { cra_set_error("tensor-memory CCF kernel: maxrin too small"); return 1; }
    int h_res[4][8]; memset(h_res, 0, sizeof(h_res));
    std::vector<int> ridx(N2, -1), rquad(N2, -1);
    for (int q = 0; q < 4; ++q) {
        if ((int)qres[q].size() != RQ) { cra_set_error("tensor-memory CCF kernel: residue assignment failed"); return 1; }
        for (int i = 0; i < RQ; ++i) { h_res[q][i] = qres[q][i]; ridx[qres[q][i]] = i; rquad[qres[q][i]] = q; }
    }
    // frequencies of quadrant q -> its warps q, q + 4, ... (longest-processing-time first)
    std::vector<std::vector<int>> lists(kWarps), flush(kWarps);
    std::vector<int> load(kWarps, 0);
    int maxc = 0;
    for (int k = 0; k < nk; ++k) maxc = std::max(maxc, nch(k));
    for (int c = maxc; c >= 1; --c)
        for (int k = 0; k < nk; ++k) {
            if (nch(k) != c) continue;
            const int n2 = k & (N2 - 1), q = rquad[n2];
            int w = q;
            for (int ww = q + 4; ww < kWarps; ww += 4) if (load[ww] < load[w]) w = ww;
            for (int j = 0; j < c; ++j) lists[w].push_back((koff[k] + j) | ((j == c - 1) ? (1 << 23) : 0));
            const int kk = (N - k) & (N - 1);
            const int col0 = (ridx[n2] * N1 + (k >> L2)) * 2;
            const int col1 = (ridx[kk & (N2 - 1)] * N1 + (kk >> L2)) * 2;
            flush[w].push_back(col0 | (col1 << 12) | ((k != 0 && k != N / 2) ? (1 << 24) : 0));
            load[w] += c;
        }
    static int h_items[kWarps][kMaxItems]; int h_n[kWarps];
    static int h_flush[kWarps][kMaxFreq];
    memset(h_items, 0, sizeof(h_items)); memset(h_flush, 0, sizeof(h_flush));
    g_sched.istride = 0; g_sched.fstride = 0;
    for (int w = 0; w < kWarps; ++w) {
        if ((int)lists[w].size() > kMaxItems || (int)flush[w].size() > kMaxFreq || koff[nk] > 8191) {
            cra_set_error("ring table too large for the tensor-memory CCF schedule"); return 1;
        }
        h_n[w] = (int)lists[w].size();
        for (size_t i = 0; i < lists[w].size(); ++i) h_items[w][i] = lists[w][i];
        for (size_t i = 0; i < flush[w].size(); ++i) h_flush[w][i] = flush[w][i];
        g_sched.istride = std::max(g_sched.istride, (int)lists[w].size());
        g_sched.fstride = std::max(g_sched.fstride, (int)flush[w].size());
    }
    CRA_CUDA(cudaStreamSynchronize(st));
    CRA_CUDA(cudaMemcpyToSymbol(g_items, h_items, sizeof(h_items)));
    CRA_CUDA(cudaMemcpyToSymbol(g_nitems, h_n, sizeof(h_n)));
    CRA_CUDA(cudaMemcpyToSymbol(g_flush, h_flush, sizeof(h_flush)));
    CRA_CUDA(cudaMemcpyToSymbol(g_res, h_res, sizeof(h_res)));
    g_sched.nring = h.nring; g_sched.maxrin = h.maxrin; g_sched.dev = dev; g_sched.len = len;
    return 0;
}

template <int LOG2N>
int launch_t(const unsigned char* spec, int nrows, const unsigned char* refspec, int R, size_t row_bytes,
             const float2* twid, CraCand* cand, int ntile_n, const float2* norm, const float* tref, cudaStream_t st)
{
    using S = TShape<LOG2N>;
    const size_t smem = ((size_t)32 * S::PS + S::N) * sizeof(float2) + (size_t)kWarps * (g_sched.istride + g_sched.fstride) * sizeof(int);
    static size_t configured = 0;
    if (smem > configured) {
        CRA_CUDA(cudaFuncSetAttribute(ccf_tm_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int nquad = (R + 3) / 4;
    const long ncta_m = (nrows + 7) / 8;
    const long nblk = ncta_m * ntile_n;
    if (nblk <= 0) return 0;
    if (nblk > 2147483647L) { cra_set_error("ccf grid too large; lower row_batch"); return 1; }
    static int resident = 0;                     // CTAs the device holds at once (2 per SM: TMEM and shared memory)
    if (!resident) {
        int dev = 0, nsm = 0, per = 0;
        CRA_CUDA(cudaGetDevice(&dev));
        CRA_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
        // CTAs one SM holds: registers, shared memory (+1 KB reserved per CTA) and the 512 TMEM columns.  The
        // occupancy API answers 1 for this kernel at 128 registers although two CTAs are co-resident
        // (measured: a 2-per-SM persistent grid runs at the two-CTA rate), so the limits are taken directly.
        CRA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, ccf_tm_kernel<LOG2N>, kThreads, smem));
        cudaFuncAttributes fa; CRA_CUDA(cudaFuncGetAttributes(&fa, ccf_tm_kernel<LOG2N>));
        int regs_sm = 0, smem_sm = 0;
        CRA_CUDA(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev));
        CRA_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        const int by_regs = regs_sm / (((fa.numRegs + 7) / 8 * 8) * kThreads);
        const int by_smem = (int)((size_t)smem_sm / (smem + fa.sharedSizeBytes + 1024));
        per = std::max(per, std::min(std::min(by_regs, by_smem), 2));
        if (getenv("CRA_TM_DEBUG")) fprintf(stderr, "ccf_tm: CTAs per SM %d (regs %d -> %d, smem -> %d)\n", per, fa.numRegs, by_regs, by_smem);
        if (getenv("CRA_TM_PER")) per = atoi(getenv("CRA_TM_PER"));
        if (per < 1) per = 1;
        if (per * S::COLS > 512) per = 512 / S::COLS;      // tensor memory: 512 columns per SM
        resident = nsm * per;
    }
    const char* np = getenv("CRA_TM_PERSIST");
    const long grid = (np && np[0] == '0') ? nblk : std::min<long>(nblk, resident);
    if (getenv("CRA_TM_DEBUG")) fprintf(stderr, "ccf_tm: tiles %ld grid %ld resident %d smem %zu\n", nblk, grid, resident, smem);
    ccf_tm_kernel<LOG2N><<<(unsigned)grid, kThreads, smem, st>>>(spec, nrows, refspec, R, row_bytes, twid, cand, nquad, ntile_n,
                                                                norm, tref, g_sched.istride, g_sched.fstride, nblk);
    CRA_CUDA(cudaGetLastError());
    return 0;
}

template <int LOG2N> int num_tiles_t(int R) { const int nq = (R + 3) / 4; return (nq + TShape<LOG2N>::NJ - 1) / TShape<LOG2N>::NJ; }

}  // namespace

bool cra_ccf_tm_supported(int log2n) { return log2n >= 5 && log2n <= 9; }

int cra_ccf_tm_num_tiles(int R, int log2n)
{
    switch (log2n) {
        case 5: return num_tiles_t<5>(R);   case 6: return num_tiles_t<6>(R);   case 7: return num_tiles_t<7>(R);
        case 8: return num_tiles_t<8>(R);   default: return num_tiles_t<9>(R);
    }
}

int cra_launch_ccf_tm(const unsigned char* spec, int nrows, const unsigned char* refspec, int R, const CraRingTab& htab,
                      const CraFragTab& frag, const std::vector<int>& h_koff, const float2* twid, CraCand* cand,
                      int ntile_n, const float2* norm, const float* tref, cudaStream_t st)
{
    if (bind_schedule(htab, h_koff, st)) return 1;
    const size_t rb = cra_frag_row_bytes(frag.nch);
    switch (htab.log2n) {
        case 5:  return launch_t<5>(spec, nrows, refspec, R, rb, twid, cand, ntile_n, norm, tref, st);
        case 6:  return launch_t<6>(spec, nrows, refspec, R, rb, twid, cand, ntile_n, norm, tref, st);
        case 7:  return launch_t<7>(spec, nrows, refspec, R, rb, twid, cand, ntile_n, norm, tref, st);
        case 8:  return launch_t<8>(spec, nrows, refspec, R, rb, twid, cand, ntile_n, norm, tref, st);
        case 9:  return launch_t<9>(spec, nrows, refspec, R, rb, twid, cand, ntile_n, norm, tref, st);
        default: cra_set_error("tensor-memory CCF kernel: maxrin must be a power of two in [32, 512]"); return 1;
    }
}
