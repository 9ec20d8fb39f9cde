// cra_ccf_tm.cu -- Crosrng_ms ring contraction on the tensor cores with the correlation spectrum
// staged in TENSOR MEMORY, fused with the inverse FFT and the peak search (EMAN2 Util::Crosrng_ms +
// the best-of loop of Util::multiref_polar_ali_2d; reference call site test_mref.py:200-201;
// replaces cu_ccf_mult_m + cuFFT C2R + cu_max_idx_batch, cuda/gpu_aln_noref.cu:1009-1143,
// :2198-2206, :1305-1346).
//
// Arithmetic: mma.sync.m16n8k16 split-bf16 x3 (a_hi b_lo + a_lo b_hi + a_hi b_hi, FP32 accumulate),
// W = q + i t Hermitian-extended, one complex inverse FFT per (row, reference) pair, maximum of q and
// of t.  Where W lives: a warp writes its finished frequencies to TMEM with tcgen05.st: in the mma
// accumulator layout lane (g, t) owns the pairs (row g, reference 4j + t) in EVERY warp, and a thread
// may only touch the TMEM lanes of its warp's quadrant (warp % 4), so the frequencies are dealt to the
// quadrants by residue n2 = k mod N2 (and N2 - n2: the Hermitian partner W[N-k] of the same MMA
// result).  Pass 1 of the inverse FFT (N1-point DFTs over k = n1 N2 + n2 at fixed n2) then reads
// nothing but the thread's own TMEM lane (tcgen05.ld .32x32b), and only its output goes to shared
// memory, one 4-reference batch (32 pairs) at a time, for pass 2 and the maxima.
//
// maxrin <= 256: a CTA is 8 warps, 8 rows x 8 references, 256 TMEM columns and ~76 KB of shared memory:
// TWO CTAs are resident per SM and the contraction of one runs under the inverse FFT of the other.
// maxrin = 512: tensor memory holds 64 pairs (one pair = 1024 words), so there is one CTA per SM; it
// runs 16 warps (4 per quadrant) to double the operand loads in flight.
//
// Round 2: (1) the kernel keeps only the maximum VALUE of q and of t per pair (FMNMX) -- the lag of a
// particle's winner is re-derived in double precision by finalize_kernel from the spectra, as EMAN2's
// own double-precision Crosrng_ms would; (2) the per-warp work list is one stream of (chunk offset, flush
// descriptor) pairs read with 128-bit shared loads and padded to whole groups of four with the all-zero
// chunk of the fragment layout (cra_common.cuh), so the inner loop has no tail conditionals and spends one
// hardware scoreboard on its bookkeeping instead of five, which leaves scoreboards for the operand loads of
// the four register sets (scripts/sass_ctrl.py prints the assignment); (3) the work lists live in
// per-context device buffers, not in __device__ globals.
#include "cra_common.cuh"
#include "cra_fft.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <algorithm>

#include <mutex>
#include <set>

namespace {

constexpr int kMaxWarps = 16;
#ifndef CRA_TM_CONST_TW
#define CRA_TM_CONST_TW 1
#endif
#ifndef CRA_TM_P2_UNROLL
#define CRA_TM_P2_UNROLL 1          // pass 2: both items of a thread in one body (3.87 -> 3.82 ms per 5.0M alignments)
#endif
#ifndef CRA_TM_LAZY
#define CRA_TM_LAZY 1               // a tile ends without a barrier (config 4: 9.76 -> 9.58 ms per 5.0M alignments; config 2 unchanged)
#endif
#ifndef CRA_TM_EXP_NOLOAD
#define CRA_TM_EXP_NOLOAD 0
#endif
#ifndef CRA_TM_EXP                  // timing experiments (not valid kernels): 2 no inverse FFT, 4 no contraction, 16 no pass 2, 32 no pass 1
#define CRA_TM_EXP 0
#endif
#ifndef CRA_TM_P1_PIPE
#define CRA_TM_P1_PIPE 1            // pass 1: the second residue's tensor-memory load in flight under the first transform (3.81 -> 3.77)
#endif

// Pass-1 twiddles exp(+2 pi i n2 k1 / N) of every supported N (32 .. 1024), table of N at offset N - 32.  The index of a
// load is uniform over the warp (n2 is the warp's residue, k1 a constant), so the constant cache serves it without an
// LSU wavefront -- the shared-memory copy cost 15 broadcast wavefronts per 16-point transform on the pipe that the
// operand loads of the co-resident CTA queue on.  Content depends on N only: written once per device.
__constant__ float2 c_itw[2048];

using crafft::fft_reg;

template <int LOG2N>
struct TShape {
    static constexpr int N = 1 << LOG2N;
    static constexpr int L1 = LOG2N / 2;
    static constexpr int L2 = LOG2N - L1;
    static constexpr int N1 = 1 << L1;
    static constexpr int N2 = 1 << L2;
    static constexpr int PS = N1 * (N2 + 1) + ((N1 * (N2 + 1)) % 2 == 0 ? 1 : 0);   // odd float2 stride of one pair
    static constexpr int NJ = 2;                        // reference quads per CTA
    static constexpr int RQ = N2 / 4;                   // residues per quadrant
    static constexpr int JCOLS = N / 2;                 // TMEM columns of one pair slot in one quadrant = RQ * N1 * 2
    static constexpr int COLS = (NJ * JCOLS < 32) ? 32 : NJ * JCOLS;
    static constexpr int KW = (LOG2N >= 9) ? 16 : 8;    // warps per CTA (a multiple of 4: KW / 4 share a TMEM quadrant)
    static constexpr int PER_SM = (COLS <= 256) ? 2 : 1;
};

__device__ __forceinline__ void mma_bf16(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// experiments (profiles/README.md): operand loads that do not allocate in L1 (its hit rate is 6 %)
#ifndef CRA_TM_LD_NOALLOC
#define CRA_TM_LD_NOALLOC 0
#endif
#if CRA_TM_LD_NOALLOC
#define CRA_TM_LDQ "ld.global.nc.L1::no_allocate"
#else
#define CRA_TM_LDQ "ld.global.nc"
#endif
struct Frag8 { unsigned w[8]; };
__device__ __forceinline__ Frag8 ldg256(const unsigned char* p)
{
    Frag8 f;
    asm volatile(CRA_TM_LDQ ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(f.w[0]), "=r"(f.w[1]), "=r"(f.w[2]), "=r"(f.w[3]), "=r"(f.w[4]), "=r"(f.w[5]), "=r"(f.w[6]), "=r"(f.w[7])
                 : "l"(p));
    return f;
}
__device__ __forceinline__ uint4 ldg128(const unsigned char* p)
{
    uint4 v;
    asm volatile(CRA_TM_LDQ ".v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
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

// the same load split into issue and completion, so that the next residue's load is in flight under a transform:
// the wait names the registers as read-write operands, which keeps every use behind it
template <int NC> struct TmRegs { unsigned r[NC]; };
__device__ __forceinline__ void tmem_issue(unsigned taddr, TmRegs<8>& v)
{
    unsigned* r = v.r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_issue(unsigned taddr, TmRegs<16>& v)
{
    unsigned* r = v.r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_issue(unsigned taddr, TmRegs<32>& v)
{
    unsigned* r = v.r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_complete(TmRegs<8>& v)
{
    unsigned* r = v.r;
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
}
__device__ __forceinline__ void tmem_complete(TmRegs<16>& v)
{
    unsigned* r = v.r;
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
}
__device__ __forceinline__ void tmem_complete(TmRegs<32>& v)
{
    unsigned* r = v.r;
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]) :: "memory");
}

struct Operands { Frag8 a; uint4 b0, b1; };

// header of the per-context schedule (device): items per warp, residues n2 of quadrant q in TMEM order
struct SchedHdr { int nit[kMaxWarps]; int res[4][8]; };

template <int LOG2N>
__global__ void __launch_bounds__(TShape<LOG2N>::KW * 32, TShape<LOG2N>::PER_SM)
ccf_tm_kernel(const unsigned char* __restrict__ spec, int nrows, const unsigned char* __restrict__ refspec, int R,
              size_t row_bytes, const float2* __restrict__ twid, CraCand* __restrict__ cand,
              int nquad, int ncta_n, const float2* __restrict__ norm, const float* __restrict__ tref,
              const int2* __restrict__ g_items, const SchedHdr* __restrict__ g_hdr, int istride, long ntiles)
{
    using S = TShape<LOG2N>;
    constexpr int N = S::N, N1 = S::N1, N2 = S::N2, PS = S::PS, NJ = S::NJ, RQ = S::RQ, JCOLS = S::JCOLS, KW = S::KW;
    constexpr int kThreads = KW * 32;
    static_assert(NJ == 2, "the operand sets below hold two reference quads");
    extern __shared__ __align__(16) float2 s_dyn[];
    float2* s_y = s_dyn;                      // 32 pairs * PS : pass-1 output of one reference quad
    float2* s_tw = s_dyn + 32 * PS;           // N : s_tw[j*N2 + n2] = exp(+2 pi i n2 j / N)
    int2* s_items = reinterpret_cast<int2*>(s_tw + N);          // KW * istride items (16-byte aligned: 32 * PS + N is even)
    __shared__ CraCand s_pair[32 * NJ];
    __shared__ unsigned s_tmem;
    __shared__ SchedHdr s_hdr;

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
    for (int i = tid; i < (int)(sizeof(SchedHdr) / sizeof(int)); i += kThreads)
        reinterpret_cast<int*>(&s_hdr)[i] = reinterpret_cast<const int*>(g_hdr)[i];
    for (int i = tid; i < KW * istride; i += kThreads) s_items[i] = g_items[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // this thread's TMEM lane: bits 31:16 lane, 15:0 column
    const unsigned tbase = s_tmem + ((unsigned)(quad * 32) << 16);

    // Persistent CTA: the work lists, the twiddles and the TMEM allocation above are set up once, then
    // the CTA walks the tiles with the grid stride.  Reference tile fastest, so that the CTAs resident
    // at any moment share the row spectra in L2.
    // The best candidate of every row of a finished tile, from the pairs pass 2 left in s_pair.  With CRA_TM_LAZY a tile
    // ends WITHOUT a barrier: a warp that has finished its share of the last pass 2 goes straight on to the operand loads
    // of the next tile (pass 2 works from shared memory only; tensor memory was drained by pass 1), and the candidates
    // of tile T are written after the first barrier of tile T + 1, which every warp reaches after its pass 2 of tile T
    // and before any warp writes s_pair again.
    auto emit_cand = [&](int tl) {
        if (tid < 8) {
            const int cm = tl / ncta_n, cn = tl - cm * ncta_n;
            const int nj = qbase + (cn < qrem ? 1 : 0);
            const int row = cm * 8 + tid;
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
    };
    int prev_tile = -1;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < (int)ntiles; tile += gridDim.x) {
    {
    const int cm = tile / ncta_n, cn = tile - cm * ncta_n;
    const int nj = qbase + (cn < qrem ? 1 : 0);                 // reference quads of this tile (<= NJ)
    const int q0 = cn * qbase + min(cn, qrem);
    const int row0 = (int)(cm * 8);
    // the previous tile's tcgen05.ld (pass 1) were ordered before its barriers
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- contraction: this warp's frequencies, chunk by chunk ---------------------------------
    if (!(CRA_TM_EXP & 4)) {
        int rrow = row0 + g;
        if (rrow >= nrows) rrow = nrows - 1;
        const unsigned char* pa = spec + (size_t)rrow * row_bytes + t * 32;
        const unsigned char* pb0 = refspec + (size_t)(4 * q0 + (g >> 1)) * row_bytes + t * 32 + (g & 1) * 16;
        const unsigned char* pb1 = refspec + (size_t)(4 * (q0 + (nj > 1 ? 1 : 0)) + (g >> 1)) * row_bytes + t * 32 + (g & 1) * 16;
        // keep the three operand bases in registers: re-deriving them from the parameter bank inside the loop
        // costs an LDC scoreboard per base (see the header)
        asm volatile("" : "+l"(pa), "+l"(pb0), "+l"(pb1));
        float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};

        const int nit = s_hdr.nit[warp];                        // a multiple of 4, >= 4
        const int4* it4 = reinterpret_cast<const int4*>(s_items + warp * istride);   // (off, flush) x 2 per int4

        // a tile whose second reference quad is empty (the last tile of R = 50, every tile of R <= 4: reference-free
        // and class-bound alignment) neither loads nor multiplies it
        const bool two = nj > 1;
#define CRA_LOAD_OPS_(O, off)                                                            \
        { O.a = ldg256(pa + (unsigned)(off)); O.b0 = ldg128(pb0 + (unsigned)(off)); if (two) O.b1 = ldg128(pb1 + (unsigned)(off)); }
#if CRA_TM_EXP_NOLOAD      /* timing experiment, not a valid kernel: the steady-state loads are skipped, operands reused */
#define CRA_LOAD_OPS(O, off) { if ((off) == 0x7fffffff) CRA_LOAD_OPS_(O, off) }
#else
#define CRA_LOAD_OPS(O, off) CRA_LOAD_OPS_(O, off)
#endif
#define CRA_COMPUTE(O, fl)                                                               \
        { mma_bf16(acc0, O.a.w[0], O.a.w[1], O.a.w[2], O.a.w[3], O.b0.z, O.b0.w);   /* a_hi b_lo */ \
          if (two) mma_bf16(acc1, O.a.w[0], O.a.w[1], O.a.w[2], O.a.w[3], O.b1.z, O.b1.w);       \
          mma_bf16(acc0, O.a.w[4], O.a.w[5], O.a.w[6], O.a.w[7], O.b0.x, O.b0.y);   /* a_lo b_hi */ \
          if (two) mma_bf16(acc1, O.a.w[4], O.a.w[5], O.a.w[6], O.a.w[7], O.b1.x, O.b1.y);       \
          mma_bf16(acc0, O.a.w[0], O.a.w[1], O.a.w[2], O.a.w[3], O.b0.x, O.b0.y);   /* a_hi b_hi */ \
          if (two) mma_bf16(acc1, O.a.w[0], O.a.w[1], O.a.w[2], O.a.w[3], O.b1.x, O.b1.y);       \
          if (fl) flush_freq(fl); }

        // a finished frequency: W[k] and its Hermitian partner W[N-k] to this lane's TMEM columns.  For k = 0 and
        // k = N/2 both columns coincide and both values are equal (the imaginary parts of those bins are exact zeros).
        auto flush_freq = [&](int f) {
            const unsigned c0 = f & 4095, c1 = (f >> 12) & 4095;
            {   // c0 = A (re.re), c1 = D (row re . ref im), c2 = C (row im . ref re), c3 = B (im.im)
                const float A = acc0[0], D = acc0[1], C = acc0[2], B = acc0[3];
                // s = (A+B, A-B), tv = (C+D, D-C);  W[k] = s + tv,  W[N-k] = s - tv
                const float sx = A + B, sy = A - B, tx = C + D, ty = D - C;
                tmem_st2(tbase + c0, sx + tx, sy + ty);
                tmem_st2(tbase + c1, sx - tx, sy - ty);
                acc0[0] = acc0[1] = acc0[2] = acc0[3] = 0.f;
            }
            if (two) {
                const float A = acc1[0], D = acc1[1], C = acc1[2], B = acc1[3];
                const float sx = A + B, sy = A - B, tx = C + D, ty = D - C;
                tmem_st2(tbase + JCOLS + c0, sx + tx, sy + ty);
                tmem_st2(tbase + JCOLS + c1, sx - tx, sy - ty);
                acc1[0] = acc1[1] = acc1[2] = acc1[3] = 0.f;
            }
        };

        // Four operand sets, one chunk each, every set reloaded right after its chunk was multiplied: a load
        // has the multiplication of three other chunks to land in.
        Operands o0, o1, o2, o3;
        int4 ia = it4[0], ib = it4[1];
        CRA_LOAD_OPS_(o0, ia.x); CRA_LOAD_OPS_(o1, ia.z); CRA_LOAD_OPS_(o2, ib.x); CRA_LOAD_OPS_(o3, ib.z);
#pragma unroll 1
        for (int i = 4; ; i += 4) {
            const bool more = i < nit;
            const int f0 = ia.y, f1 = ia.w, f2 = ib.y, f3 = ib.w;
            if (more) { ia = it4[i >> 1]; ib = it4[(i >> 1) + 1]; }
            CRA_COMPUTE(o0, f0);
            if (more) CRA_LOAD_OPS(o0, ia.x);
            CRA_COMPUTE(o1, f1);
            if (more) CRA_LOAD_OPS(o1, ia.z);
            CRA_COMPUTE(o2, f2);
            if (more) CRA_LOAD_OPS(o2, ib.x);
            CRA_COMPUTE(o3, f3);
            if (!more) break;
            CRA_LOAD_OPS(o3, ib.z);
        }
#undef CRA_LOAD_OPS
#undef CRA_LOAD_OPS_
#undef CRA_COMPUTE
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#if CRA_TM_LAZY
    if (prev_tile >= 0) emit_cand(prev_tile);
#endif
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
        if (CRA_TM_EXP & 2) break;
        // pass 1: (pair = lane, residue n2): N1-point DFT over n1 straight from this lane's TMEM, twiddle
        if (!(CRA_TM_EXP & 32)) {
        auto p1_emit = [&](float2 (&x)[N1], int n2) {            // twiddle and park the transform of residue n2
            float2* w = s_y + lane * PS + n2;
#pragma unroll
            for (int k1 = 0; k1 < N1; ++k1) {
                if (k1 == 0) { w[0] = x[0]; continue; }
                const float2 tw = (CRA_TM_CONST_TW && LOG2N <= 8) ? c_itw[(N - 32) + k1 * N2 + n2] : s_tw[k1 * N2 + n2];   // 4 KB at N = 512 overflow the constant cache: +4 % there
                w[k1 * (N2 + 1)] = crafft::cmul(x[k1], tw);
            }
        };
        constexpr int P1IT = (RQ + KW / 4 - 1) / (KW / 4);       // residues per warp (1 or 2)
#if CRA_TM_P1_PIPE
        {
        TmRegs<2 * N1> tr[2];
        tmem_issue(tbase + j * JCOLS + (warp >> 2) * (2 * N1), tr[0]);
#pragma unroll
        for (int pit = 0; pit < P1IT; ++pit) {
            const int ri = (warp >> 2) + pit * (KW / 4);
            if (ri >= RQ) break;
            float2 x[N1];
            tmem_complete(tr[pit & 1]);
            if (pit + 1 < P1IT && ri + KW / 4 < RQ) tmem_issue(tbase + j * JCOLS + (ri + KW / 4) * (2 * N1), tr[(pit + 1) & 1]);
#pragma unroll
            for (int i = 0; i < N1; ++i) x[i] = make_float2(__uint_as_float(tr[pit & 1].r[2 * i]), __uint_as_float(tr[pit & 1].r[2 * i + 1]));
            fft_reg<N1, 1>(x);
            p1_emit(x, s_hdr.res[quad][ri]);
        }
        }
#else
        for (int ri = warp >> 2; ri < RQ; ri += KW / 4) {
            float2 x[N1];
            tmem_ld<2 * N1>(tbase + j * JCOLS + ri * (2 * N1), x);
            fft_reg<N1, 1>(x);
            p1_emit(x, s_hdr.res[quad][ri]);
        }
#endif
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // pass 2: (pair, k1): N2-point DFT over n2 -> X[k1 + N1*k2]; only the MAXIMA of q = Re X and t = Im X are
        // kept -- finalize_kernel re-derives the lag of the particle's winner in double precision
#if CRA_TM_P2_UNROLL
#pragma unroll 2
#endif
        for (int item = tid; item < 32 * N1; item += kThreads) {
            if (CRA_TM_EXP & 16) break;
            const int pi = item / N1, k1 = item - pi * N1;
            const float2* w = s_y + pi * PS + k1 * (N2 + 1);
            float2 x[N2];
#pragma unroll
            for (int jj = 0; jj < N2; ++jj) x[jj] = w[jj];
            fft_reg<N2, 1>(x);
            float bq = x[0].x, bt = x[0].y;
#pragma unroll
            for (int jj = 1; jj < N2; ++jj) { bq = fmaxf(bq, x[jj].x); bt = fmaxf(bt, x[jj].y); }
            // the N1 lanes of one pair are consecutive and aligned inside a warp (N1 <= 32)
#pragma unroll
            for (int o = N1 >> 1; o > 0; o >>= 1) {
                bq = fmaxf(bq, __shfl_xor_sync(0xffffffffu, bq, o));
                bt = fmaxf(bt, __shfl_xor_sync(0xffffffffu, bt, o));
            }
            if (k1 == 0) {
                const int row = row0 + (pi >> 2), ref = 4 * (q0 + j) + (pi & 3);     // pair = lane (g, t)
                CraCand cd;
                if (row < nrows && ref < R) {
                    // deferred Normalize_ring (cra_common.cuh): every lag moves by -avg * tref[ref]
                    const float2 nm = norm[row];
                    const float sc = nm.y / (float)N, dc = nm.x * tref[ref];
                    const float qn = (bq - dc) * sc, qm = (bt - dc) * sc;
                    if (qn >= qm) { cd.v = qn; cd.code = ref * 8192; }
                    else          { cd.v = qm; cd.code = ref * 8192 + 4096; }
                } else { cd.v = -INFINITY; cd.code = -1; }
                s_pair[j * 32 + pi] = cd;
            }
        }
#if CRA_TM_LAZY
        if (j + 1 < nj) __syncthreads();         // the next quad's pass 1 overwrites s_y
#else
        __syncthreads();
#endif
    }
#if CRA_TM_LAZY
    prev_tile = tile;
#else
    (void)cn;
    emit_cand(tile);
#endif
    }
    }   // tile loop
#if CRA_TM_LAZY
    __syncthreads();
    if (prev_tile >= 0) emit_cand(prev_tile);
#endif
    // every tcgen05.ld has completed (wait::ld) before the barriers above
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem), "n"(S::COLS) : "memory");
    }
}

// ---- per-context schedule ------------------------------------------------------------------------------
struct TmSched {
    int2* d_items = nullptr;
    SchedHdr* d_hdr = nullptr;
    int istride = 0;           // items per warp (a multiple of 4)
    int kw = 0;
    int resident = 0;          // CTAs the device holds at once
};

// Deal the frequencies to the TMEM quadrants by residue n2 = k mod N2 (a residue and its negative
// together), balance each quadrant's frequencies over its warps, and upload the lists.
int build_schedule(const CraRingTab& h, const std::vector<int>& koff, int kw, cudaStream_t st, TmSched** out)
{
    const int N = h.maxrin, nk = N / 2 + 1;
    const int L1 = h.log2n / 2, L2 = h.log2n - L1, N1 = 1 << L1, N2 = 1 << L2, RQ = N2 / 4;
    if (RQ < 1 || RQ > 8) { cra_set_error("tensor-memory CCF kernel: unsupported maxrin"); return 1; }
    if (N2 == 4) { cra_set_error("tensor-memory CCF kernel: maxrin too small"); return 1; }
    auto nch = [&](int k) { return koff[k + 1] - koff[k]; };
    const int zero_chunk = koff[nk];                                   // chunk nch of every row: all zeros
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
    SchedHdr hdr; memset(&hdr, 0, sizeof(hdr));
    std::vector<int> ridx(N2, -1), rquad(N2, -1);
    for (int q = 0; q < 4; ++q) {
        if ((int)qres[q].size() != RQ) { cra_set_error("tensor-memory CCF kernel: residue assignment failed"); return 1; }
        for (int i = 0; i < RQ; ++i) { hdr.res[q][i] = qres[q][i]; ridx[qres[q][i]] = i; rquad[qres[q][i]] = q; }
    }
    // frequencies of quadrant q -> its warps q, q + 4, ... (longest-processing-time first)
    std::vector<std::vector<int2>> lists(kw);
    std::vector<int> load(kw, 0);
    int maxc = 0;
    for (int k = 0; k < nk; ++k) maxc = std::max(maxc, nch(k));
    if ((size_t)(zero_chunk + 1) * 128 >= ((size_t)1 << 31)) { cra_set_error("ring table too large for the tensor-memory CCF schedule"); return 1; }
    for (int c = maxc; c >= 1; --c)
        for (int k = 0; k < nk; ++k) {
            if (nch(k) != c) continue;
            const int n2 = k & (N2 - 1), q = rquad[n2];
            int w = q;
            for (int ww = q + 4; ww < kw; ww += 4) if (load[ww] < load[w]) w = ww;
            const int kk = (N - k) & (N - 1);
            const int col0 = (ridx[n2] * N1 + (k >> L2)) * 2;
            const int col1 = (ridx[kk & (N2 - 1)] * N1 + (kk >> L2)) * 2;
            const int fl = col0 | (col1 << 12) | (1 << 24);
            for (int j = 0; j < c; ++j) lists[w].push_back(make_int2((koff[k] + j) * 128, (j == c - 1) ? fl : 0));
            load[w] += c;
        }
    TmSched* s = new TmSched();
    s->kw = kw;
    for (int w = 0; w < kw; ++w) {
        while (lists[w].size() < 4 || (lists[w].size() & 3)) lists[w].push_back(make_int2(zero_chunk * 128, 0));
        hdr.nit[w] = (int)lists[w].size();
        s->istride = std::max(s->istride, (int)lists[w].size());
    }
    std::vector<int2> flat((size_t)kw * s->istride, make_int2(zero_chunk * 128, 0));
    for (int w = 0; w < kw; ++w) std::copy(lists[w].begin(), lists[w].end(), flat.begin() + (size_t)w * s->istride);
    cudaError_t e = cudaMalloc(&s->d_items, flat.size() * sizeof(int2));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_hdr, sizeof(SchedHdr));
    // stream-ordered copies (pageable source: staged by the runtime before the call returns), ordered before the kernels on st
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_items, flat.data(), flat.size() * sizeof(int2), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_hdr, &hdr, sizeof(SchedHdr), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cra_set_error(std::string("tensor-memory CCF schedule: ") + cudaGetErrorString(e));
        cudaFree(s->d_items); cudaFree(s->d_hdr); delete s; return 1;
    }
    *out = s;
    return 0;
}

int ensure_const_twiddles(cudaStream_t st)
{
    static std::mutex mu;
    static std::set<int> done;
    static std::vector<float2> host;                 // stays alive: the copies below are asynchronous
    int dev = 0;
    CRA_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (done.count(dev)) return 0;
    if (host.empty()) {
        host.assign(2048, make_float2(0.f, 0.f));
        for (int lg = 5; lg <= 10; ++lg) {
            std::vector<float2> tw; cra_ccf_twiddles(lg, tw);
            std::copy(tw.begin(), tw.end(), host.begin() + ((1 << lg) - 32));
        }
    }
    CRA_CUDA(cudaMemcpyToSymbolAsync(c_itw, host.data(), host.size() * sizeof(float2), 0, cudaMemcpyHostToDevice, st));
    CRA_CUDA(cudaStreamSynchronize(st));
    done.insert(dev);
    return 0;
}

template <int LOG2N>
int launch_t(const unsigned char* spec, int nrows, const unsigned char* refspec, int R, size_t row_bytes,
             const float2* twid, CraCand* cand, int ntile_n, const float2* norm, const float* tref,
             const CraRingTab& htab, const std::vector<int>& h_koff, void** slot, cudaStream_t st)
{
    using S = TShape<LOG2N>;
    constexpr int kThreads = S::KW * 32;
    TmSched* sc = static_cast<TmSched*>(*slot);
    if (!sc) {
        if (build_schedule(htab, h_koff, S::KW, st, &sc)) return 1;
        *slot = sc;
    }
    const size_t smem = ((size_t)32 * S::PS + S::N) * sizeof(float2) + (size_t)S::KW * sc->istride * sizeof(int2);
    if (cra_ensure_dyn_smem(reinterpret_cast<const void*>(&ccf_tm_kernel<LOG2N>), smem)) return 1;
    if (CRA_TM_CONST_TW && ensure_const_twiddles(st)) return 1;
    const int nquad = (R + 3) / 4;
    const long ncta_m = (nrows + 7) / 8;
    const long nblk = ncta_m * ntile_n;
    if (nblk <= 0) return 0;
    if (nblk > 2147483647L) { cra_set_error("ccf grid too large; lower row_batch"); return 1; }
    if (!sc->resident) {                         // CTAs the device holds at once (2 per SM: TMEM and shared memory)
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
        per = std::max(per, std::min(std::min(by_regs, by_smem), S::PER_SM));
        if (getenv("CRA_TM_DEBUG")) fprintf(stderr, "ccf_tm: CTAs per SM %d (regs %d -> %d, smem -> %d)\n", per, fa.numRegs, by_regs, by_smem);
        if (getenv("CRA_TM_PER")) per = atoi(getenv("CRA_TM_PER"));
        if (per < 1) per = 1;
        if (per * S::COLS > 512) per = 512 / S::COLS;      // tensor memory: 512 columns per SM
        sc->resident = nsm * per;
    }
    const long grid = std::min<long>(nblk, sc->resident);
    if (getenv("CRA_TM_DEBUG")) fprintf(stderr, "ccf_tm: tiles %ld grid %ld resident %d smem %zu\n", nblk, grid, sc->resident, smem);
    ccf_tm_kernel<LOG2N><<<(unsigned)grid, kThreads, smem, st>>>(spec, nrows, refspec, R, row_bytes, twid, cand, nquad, ntile_n,
                                                                norm, tref, sc->d_items, sc->d_hdr, sc->istride, nblk);
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

void cra_ccf_tm_sched_free(void* sched)
{
    TmSched* s = static_cast<TmSched*>(sched);
    if (!s) return;
    cudaFree(s->d_items); cudaFree(s->d_hdr);
    delete s;
}

int cra_launch_ccf_tm(const unsigned char* spec, int nrows, const unsigned char* refspec, int R, const CraRingTab& htab,
                      const CraFragTab& frag, const std::vector<int>& h_koff, const float2* twid, CraCand* cand,
                      int ntile_n, const float2* norm, const float* tref, void** sched, cudaStream_t st)
{
    const size_t rb = cra_frag_row_bytes(frag.nch);
    switch (htab.log2n) {
        case 5:  return launch_t<5>(spec, nrows, refspec, R, rb, twid, cand, ntile_n, norm, tref, htab, h_koff, sched, st);
        case 6:  return launch_t<6>(spec, nrows, refspec, R, rb, twid, cand, ntile_n, norm, tref, htab, h_koff, sched, st);
        case 7:  return launch_t<7>(spec, nrows, refspec, R, rb, twid, cand, ntile_n, norm, tref, htab, h_koff, sched, st);
        case 8:  return launch_t<8>(spec, nrows, refspec, R, rb, twid, cand, ntile_n, norm, tref, htab, h_koff, sched, st);
        case 9:  return launch_t<9>(spec, nrows, refspec, R, rb, twid, cand, ntile_n, norm, tref, htab, h_koff, sched, st);
        default: cra_set_error("tensor-memory CCF kernel: maxrin must be a power of two in [32, 512]"); return 1;
    }
}
