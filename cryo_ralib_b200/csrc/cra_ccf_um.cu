// cra_ccf_um.cu -- Crosrng_ms ring contraction on tcgen05.mma (UMMA) with the correlation spectrum
// streamed through TENSOR MEMORY one residue class at a time, fused with the inverse FFT and the
// peak search (EMAN2 Util::Crosrng_ms + the best-of loop of Util::multiref_polar_ali_2d; reference
// call site test_mref.py:200-201; replaces cu_ccf_mult_m + cuFFT C2R + cu_max_idx_batch,
// cuda/gpu_aln_noref.cu:1009-1143, :2198-2206, :1305-1346).
//
// Why this shape.  One (row, reference) pair needs its whole correlation spectrum W (2 KB) on chip
// for the inverse FFT, so an SM can hold ~128 pairs, while a tcgen05.mma tile is M >= 64 lanes.  The
// mapping that fits: M = 64 = 32 rows x {re, im}, N = 8 = 4 references x {U, V} (U = re - im,
// V = re + im of the weighted reference spectrum), K = 16 = 8 rings x {bf16 hi, bf16 lo} of the row:
//   D[(row, re)][(ref, U)] = sum a U   D[(row, re)][(ref, V)] = sum a V      (a + i b = row spectrum)
//   D[(row, im)][(ref, U)] = sum b U   D[(row, im)][(ref, V)] = sum b V
//   W[k] = (aV + bV) + i (aV - bV),   W[N-k] = (aU - bU) + i (aU + bU)     (= q + i t of Crosrng_ms)
// Split precision costs two MMAs per 8 rings instead of three: the row operand carries hi and lo
// side by side along K, the reference operand is [c_hi | c_hi] (MMA 1: (hi + lo) c_hi) and
// [c_lo | 0] (MMA 2: hi c_lo); the dropped lo.lo term is 2^-18 relative.
// The four-step inverse FFT (N = N1 N2, k = n1 N2 + n2) needs, for pass 1 at residue n2, only the
// frequencies k = n2 (mod N2) and, by Hermitian symmetry, the partners of k = -n2: the MMA warp
// walks the frequencies CLASS by class {n2, N2 - n2}; a class is 16 frequencies x 8 columns = 128
// TMEM columns in the 16 lower (buffer 0) or upper (buffer 1) lanes of every quadrant (the M = 64
// accumulator layout), double buffered against the consumer warps.  With the m16n8-like
// tcgen05.ld.16x256b every consumer thread receives W[k] and W[N-k] of ITS pair (row 8q + lane/4,
// reference lane%4), runs the two N1-point DFTs of the class in registers and parks the result Y in
// its own TMEM lane (12 of the 16 residues, 32x32b, 384 columns) or its own shared-memory column (the
// other 4).  Pass 2 (N2-point DFTs over the residues), the ">=" argmax, the straight/mirror choice and
// the best reference are then thread-local.
// Operands.  With CRA_CCF=um the row kernels write particle rows in the reference layout
// (CraFragTab::unit_rows, cra_common.cuh): a 16-byte unit [hi x4 | lo x4] of one (row, part) is one
// row of a UMMA K-major core matrix, so a producer warp moves it with one 16-byte cp.async, no
// registers held (the CUTLASS sm100 cp.async mainloop pattern: completion arrives on the stage's
// mbarrier, no proxy fence).  The reference operand is a prebuilt 1 KB image per (4-reference tile,
// chunk).  One thread issuing every MMA paces the contraction (~55 SASS instructions per chunk), so the
// frequency slots of a class are dealt to FOUR independent pipelines (producer warp + MMA warp +
// 7-stage ring each): different slots are different accumulator columns and need no ordering.
// One CTA per SM: 8 consumer warps + 4 producer warps + 4 MMA warps, 207 KB of shared memory, all 512
// TMEM columns.
// Status (profiles/README.md): parity suite green; 10.6 ms per 5.0M alignments against 4.8 ms of the
// mma.sync kernel (cra_ccf_tm.cu), which stays the default.  ncu: tensor pipe 4.5 %, L2->SM 63 GB per
// launch = 12 KB per pair at 51 % of the L2 throughput peak: a 128-pair tile re-reads its 32 rows for
// every 4 references, and that operand delivery binds, as it does for the mma.sync kernel (34 GB).
// What would lift it: cluster multicast of the row operand (TMA 5-D tensor map box = one chunk of
// 32 rows in core-matrix order, multicast to the CTAs of a cluster that share the row tile).
// CRA_UM_DBG bits (timing experiments, not valid kernels): 1 no MMAs, 8 no pass 1, 16 one pass-2
// iteration, 32 clock64 timestamps of CTA 200 printed by the launcher.
#include "cra_common.cuh"
#include "cra_fft.cuh"
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>

namespace {

constexpr int LOG2N = 8, N = 256, N1 = 16, N2 = 16;
constexpr int kConsWarps = 8;
constexpr int kPipes = 4;                   // independent operand pipelines: one producer warp + one MMA warp each
constexpr int kThreads = (kConsWarps + 2 * kPipes) * 32;         // 512
constexpr int kStages = 7;                  // shared-memory operand stages per pipeline (one chunk each)
constexpr int kABytes = 4096, kBBytes = 1024;          // per chunk
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kMaxItems = 256;              // chunks per pipeline
constexpr int kTmemCols = 512;
constexpr int kYCol = 128;                  // first TMEM column of the Y store (D tiles: columns 0..127)
constexpr int kYT = 12;                     // residues whose Y lives in TMEM (12 x 16 x 2 = 384 columns); the other 4 in shared memory
constexpr int kPfDist = 36;                 // row tiles between an L2 prefetch and its use (~11 row tiles are resident at a time)
constexpr unsigned kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 17) | (4u << 24);   // f32 acc, bf16 x bf16, K-major, N = 8, M = 64

// work list, chunk by chunk in class order:
//   bits 0..6 TMEM column of the frequency slot (slot * 8) | 7..19 chunk index gc | 20 class buffer (TMEM lane 16)
//   | 24 first chunk of its frequency | 25 last chunk of its class | 26 first chunk of its class | 27 valid
__device__ int g_items[kPipes][kMaxItems];
__device__ int g_nitems[kPipes];
__device__ long long g_dbg[32];
#define UM_T(i) do { if ((dbg & 32) && blockIdx.x == 200) g_dbg[i] = clock64(); } while (0)

using crafft::fft_reg;
using crafft::cmul;

__device__ __forceinline__ bool better(float v, int m, float bv, int bm)
{   // ">=" scan order semantics: larger value wins, ties go to the later index
    return (v > bv) || (v == bv && m > bm);
}

// ---- mbarrier / tcgen05 helpers -------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity)
{
    const unsigned a = smem_u32(b);
    unsigned done = 0, spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();       // a lost arrival must not hang the device
    }
}
// the same for a warp that can afford to be late: a failed poll backs off, so that the consumer warps waiting for a
// class do not take the issue slots of the producer and MMA warps that share their schedulers
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long* b, unsigned parity)
{
    const unsigned a = smem_u32(b);
    unsigned done = 0, spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(256);
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(unsigned long long* b)      // arrives once this thread's copies have landed
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(unsigned long long* b)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ unsigned long long umma_desc(unsigned addr, unsigned lbo, unsigned sbo)
{   // K-major, no swizzle: core matrix = 8 rows x 16 B; LBO = byte step between the two K halves, SBO = between 8-row groups
    return (unsigned long long)((addr >> 4) & 0x3fffu) | ((unsigned long long)((lbo >> 4) & 0x3fffu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
// 8 frequency slots (64 columns) of a class tile: r[4s + {0,1,2,3}] = {aU, aV, bU, bV} of slot s for this thread's pair
__device__ __forceinline__ void tmem_ld_16x256b_x8(unsigned taddr, float (&r)[32])
{
    unsigned u[32];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
                   "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
                   "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_ld_16x256b_x1(unsigned taddr, float (&r)[4])
{
    unsigned u[4];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(unsigned taddr, float2 (&x)[8])
{
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(unsigned taddr, float2 (&x)[4])
{
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
}
__device__ __forceinline__ void tmem_st2(unsigned taddr, float2 v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(taddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)) : "memory");
}
__device__ __forceinline__ void cons_sync()      // the consumer warps only
{
    asm volatile("bar.sync 1, %0;" :: "n"(kConsWarps * 32) : "memory");
}

// Y of residue n2, output k1: TMEM column (n2 < 8) or shared-memory float2 index (n2 >= 8) of this pair
__device__ __forceinline__ void store_y(unsigned ybase, float2* sy, int n2, int k1, float2 v)
{
    if (n2 < kYT) tmem_st2(ybase + (k1 * kYT + n2) * 2, v);
    else sy[(k1 * (N2 - kYT) + (n2 - kYT)) * 128] = v;
}
// pass 1 of one residue: x[n1] = W[n1 N2 + n2] -> N1-point inverse DFT, twiddle, park
__device__ __forceinline__ void pass1(float2 (&x)[N1], int n2, unsigned ybase, float2* sy, const float2* __restrict__ s_tw)
{
    fft_reg<N1, 1>(x);
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) {
        const float2 v = (k1 == 0) ? x[0] : cmul(x[k1], s_tw[k1 * N2 + n2]);
        store_y(ybase, sy, n2, k1, v);
    }
}

__global__ void __launch_bounds__(kThreads, 1)
ccf_um_kernel(const unsigned char* __restrict__ spec, int nrows, const unsigned char* __restrict__ refimg, int R,
              size_t row_bytes, int nch, const float2* __restrict__ twid, CraCand* __restrict__ cand, int ncta_n,
              const float2* __restrict__ norm, const float* __restrict__ tref, int istride, int cls_n2_packed_lo, int cls_n2_packed_hi, int dbg)
{
    extern __shared__ __align__(1024) unsigned char s_raw[];
    unsigned char* s_stage = s_raw;                                            // kPipes * kStages * kStageBytes
    float2* s_y = reinterpret_cast<float2*>(s_raw + kPipes * kStages * kStageBytes);    // 4 * 16 * 128 float2 = 64 KB
    float2* s_tw = s_y + (N2 - kYT) * N1 * 128;                                // N
    int* s_items = reinterpret_cast<int*>(s_tw + N);                           // kPipes * istride
    __shared__ __align__(8) unsigned long long s_full[kPipes][kStages], s_empty[kPipes][kStages], s_dfull[2], s_dempty[2];
    __shared__ float4 s_merge[128];
    __shared__ unsigned s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) UM_T(0);
    const int cn = blockIdx.x % ncta_n, cm = blockIdx.x / ncta_n;
    const int row0 = cm * 32;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&s_tmem)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int p = 0; p < kPipes; ++p)
            for (int s = 0; s < kStages; ++s) { mbar_init(&s_full[p][s], 32); mbar_init(&s_empty[p][s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&s_dfull[b], kPipes); mbar_init(&s_dempty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < N; i += kThreads) s_tw[i] = twid[i];
    for (int i = tid; i < kPipes * istride; i += kThreads) s_items[i] = g_items[i / istride][i % istride];
    if (cn == 0) {                            // pull a future row tile into L2: the spectra were written a whole batch earlier
        const long r0 = (long)(cm + kPfDist) * 32;
        if (r0 < nrows) {
            const long r1 = min((long)nrows, r0 + 32);
            const unsigned char* p0 = spec + (size_t)r0 * row_bytes;
            const size_t nline = (size_t)(r1 - r0) * row_bytes / 128;
            for (size_t i = tid; i < nline; i += kThreads)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(p0 + i * 128));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem;
    if (tid == 0) UM_T(1);

    if (warp < kConsWarps) {
        // ================= consumers: pass 1 class by class, then pass 2 + argmax ====================
        const int q = warp & 3, grp = warp >> 2;
        const int pair = q * 32 + lane;                                        // row 8q + lane/4, reference lane%4
        const unsigned ybase = tmem + ((unsigned)(q * 32) << 16) + kYCol;      // own lane, 32x32b
        const unsigned dbase = tmem + ((unsigned)(q * 32 + grp * 16) << 16);   // class tile of buffer grp, 16x256b
        float2* sy = s_y + pair;
        unsigned dph = 0;
        for (int pos = grp; pos < 9; pos += 2) {
            const int n2 = ((pos < 8 ? cls_n2_packed_lo >> (4 * pos) : cls_n2_packed_hi) & 15);
            if (lane == 0) mbar_wait_relaxed(&s_dfull[grp], dph);
            dph ^= 1;
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float r[32];
            float2 xa[N1], xb[N1];
            tmem_ld_16x256b_x8(dbase, r);
            if (n2 != 0 && n2 != N2 / 2) {
                // slots 0..7: k = n2 + 16 j -> W[k] = xa[j], W[N-k] = xb[15-j];  slots 8..15: k = (16 - n2) + 16 j -> W[k] = xb[j], W[N-k] = xa[15-j]
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float aU = r[4 * j], aV = r[4 * j + 1], bU = r[4 * j + 2], bV = r[4 * j + 3];
                    xa[j] = make_float2(aV + bV, aV - bV);
                    xb[15 - j] = make_float2(aU - bU, aU + bU);
                }
                tmem_ld_16x256b_x8(dbase + 64, r);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_dempty[grp]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float aU = r[4 * j], aV = r[4 * j + 1], bU = r[4 * j + 2], bV = r[4 * j + 3];
                    xb[j] = make_float2(aV + bV, aV - bV);
                    xa[15 - j] = make_float2(aU - bU, aU + bU);
                }
                if (!(dbg & 8)) {
                pass1(xa, n2, ybase, sy, s_tw);
                pass1(xb, N2 - n2, ybase, sy, s_tw);
                }
            } else if (n2 == 0) {
                // slots 0..8: k = 16 j -> W[k] = xa[j] (j <= 8), W[N-k] = xa[16-j] (1 <= j <= 7)
                float r2[4];
                tmem_ld_16x256b_x1(dbase + 64, r2);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_dempty[grp]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float aU = r[4 * j], aV = r[4 * j + 1], bU = r[4 * j + 2], bV = r[4 * j + 3];
                    xa[j] = make_float2(aV + bV, aV - bV);
                    if (j >= 1) xa[16 - j] = make_float2(aU - bU, aU + bU);
                }
                xa[8] = make_float2(r2[1] + r2[3], r2[1] - r2[3]);
                pass1(xa, 0, ybase, sy, s_tw);
            } else {
                // slots 0..7: k = 8 + 16 j -> W[k] = xa[j], W[N-k] = xa[15-j]
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_dempty[grp]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float aU = r[4 * j], aV = r[4 * j + 1], bU = r[4 * j + 2], bV = r[4 * j + 3];
                    xa[j] = make_float2(aV + bV, aV - bV);
                    xa[15 - j] = make_float2(aU - bU, aU + bU);
                }
                pass1(xa, N2 / 2, ybase, sy, s_tw);
            }
        }
        if (tid == 0) UM_T(2);
        if (tid == 128) UM_T(3);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        cons_sync();
        if (tid == 0) UM_T(4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // pass 2: k1 = 8 grp .. 8 grp + 7 of this pair: N2-point DFT over the residues -> X[k1 + N1 k2]
        float bq = -INFINITY, bt = -INFINITY; int mq = -1, mt = -1;
#pragma unroll 1
        for (int kk = 0; kk < ((dbg & 16) ? 1 : N1 / 2); ++kk) {
            const int k1 = grp * (N1 / 2) + kk;
            float2 x[N2];
            {
                float2 lo[8], mid[4];
                tmem_ld_32x32b_x16(ybase + k1 * (2 * kYT), lo);
                tmem_ld_32x32b_x8(ybase + k1 * (2 * kYT) + 16, mid);
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = lo[j];
#pragma unroll
                for (int j = 0; j < 4; ++j) { x[8 + j] = mid[j]; x[12 + j] = sy[(k1 * (N2 - kYT) + j) * 128]; }
            }
            fft_reg<N2, 1>(x);
            float lq = -INFINITY, lt = -INFINITY; int lmq = -1, lmt = -1;
#pragma unroll
            for (int jj = 0; jj < N2; ++jj) {
                const int m = k1 + N1 * jj;
                if (x[jj].x >= lq) { lq = x[jj].x; lmq = m; }
                if (x[jj].y >= lt) { lt = x[jj].y; lmt = m; }
            }
            if (better(lq, lmq, bq, mq)) { bq = lq; mq = lmq; }
            if (better(lt, lmt, bt, mt)) { bt = lt; mt = lmt; }
        }
        if (tid == 0) UM_T(5);
        if (grp == 1) s_merge[pair] = make_float4(bq, __int_as_float(mq), bt, __int_as_float(mt));
        cons_sync();
        if (grp == 0) {
            const float4 o = s_merge[pair];
            if (better(o.x, __float_as_int(o.y), bq, mq)) { bq = o.x; mq = __float_as_int(o.y); }
            if (better(o.z, __float_as_int(o.w), bt, mt)) { bt = o.z; mt = __float_as_int(o.w); }
            const int row = row0 + 8 * q + (lane >> 2), ref = 4 * cn + (lane & 3);
            float cv = -INFINITY; int cc = -1;
            if (row < nrows && ref < R) {
                // deferred Normalize_ring (cra_common.cuh): every lag moves by -avg * tref[ref]
                const float2 nm = norm[row];
                const float sc = nm.y / (float)N, dc = nm.x * tref[ref];
                const float qn = (bq - dc) * sc, qm = (bt - dc) * sc;
                if (qn >= qm) { cv = qn; cc = ref * 8192 + (mq + 1); }
                else          { cv = qm; cc = ref * 8192 + 4096 + (mt + 1); }
            }
            // best of the tile's references in visit order: ">=", the later reference wins a tie
#pragma unroll
            for (int o2 = 1; o2 <= 2; o2 <<= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, cv, o2);
                const int oc = __shfl_xor_sync(0xffffffffu, cc, o2);
                const bool mine_low = (lane & o2) == 0;
                const float lv = mine_low ? cv : ov, hv = mine_low ? ov : cv;
                const int lc = mine_low ? cc : oc, hc = mine_low ? oc : cc;
                const bool take_hi = (hc >= 0) && (lc < 0 || hv >= lv);
                cv = take_hi ? hv : lv; cc = take_hi ? hc : lc;
            }
            if ((lane & 3) == 0 && row < nrows) {
                CraCand cd; cd.v = cv; cd.code = cc;
                cand[(size_t)row * ncta_n + cn] = cd;
            }
        }
    } else if (warp < kConsWarps + kPipes) {
        // ================= producers: fragment spectra -> UMMA core matrices ==========================
        // A 16-byte unit of the fragment layout ([hi x4 | lo x4] of one (row, part), cra_common.cuh) is one
        // core-matrix row: cp.async copies it straight to its place, no registers held, so the depth in flight
        // is the number of shared-memory stages.  Adjacent lanes read the two halves of one 32-byte sector.
        const int pipe = warp - kConsWarps;
        // one instruction = 4 rows x the 8 units of their chunk: every 8 lanes read one whole 128-byte line
        const int piece = lane & 7, j = piece >> 1, part = piece & 1;
        const unsigned char* psrc[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            int rrow = row0 + 4 * n + (lane >> 3);
            if (rrow >= nrows) rrow = nrows - 1;
            psrc[n] = spec + (size_t)rrow * row_bytes + piece * 16;
        }
        const unsigned char* pb = refimg + (size_t)cn * nch * kBBytes + lane * 16;
        const unsigned sbase = smem_u32(s_stage) + pipe * kStages * kStageBytes;
        // row rt = 4 n + lane/8: octet q = n >> 1, u8 = 4 (n & 1) + lane/8
        const unsigned dst_a = (unsigned)((j * 8 + part) * 128 + (lane >> 3) * 16);      // + (n >> 1) * 256 + (n & 1) * 64
        const unsigned dst_b = (unsigned)(kABytes + lane * 16);
        const int nit = g_nitems[pipe];
        const int* items = s_items + pipe * istride;
        unsigned eph = 1;                                       // a fresh barrier: waiting on the previous phase passes
        for (int it = 0; it < nit; ++it) {
            const int s = it % kStages;
            if (pipe == 0 && lane == 0 && it == 24) UM_T(16);
            if (lane == 0) mbar_wait(&s_empty[pipe][s], eph);
            if (pipe == 0 && lane == 0 && it == 24) UM_T(17);
            if (s == kStages - 1) eph ^= 1;
            __syncwarp();
            const unsigned st = sbase + s * kStageBytes;
            const size_t goff = (size_t)((items[it] >> 7) & 8191) * 128;
#pragma unroll
            for (int n = 0; n < 8; ++n) cp_async16(st + (n >> 1) * 256 + (n & 1) * 64 + dst_a, psrc[n] + goff);
            cp_async16(st + dst_b, pb + goff * (kBBytes / 128));
            cp_async16(st + dst_b + 512, pb + goff * (kBBytes / 128) + 512);
            cp_async_arrive(&s_full[pipe][s]);
            if (pipe == 0 && lane == 0 && it == 24) UM_T(18);
            if (pipe == 0 && lane == 0 && it == 20) UM_T(20);
        }
        if (tid == kConsWarps * 32) UM_T(11);
    } else {
        // ================= MMA issuers =================================================================
        // One thread issues the tcgen05.mma of a chunk (four of them), and its instruction stream is the pace
        // of the contraction, so the chunks of a class are dealt to kPipes warps by frequency slot: different
        // slots are different accumulator columns, no ordering between the warps is needed.
        const int pipe = warp - kConsWarps - kPipes;
        unsigned fph = 0, dph0 = 1, dph1 = 1;
        const unsigned sbase = smem_u32(s_stage) + pipe * kStages * kStageBytes;
        const unsigned long long adesc = umma_desc(sbase, 1024, 128);
        const unsigned long long bdesc = umma_desc(sbase + kABytes, 128, 256);
        const int nit = g_nitems[pipe];
        const int* items = s_items + pipe * istride;
        for (int it0 = 0; it0 < nit; it0 += kStages) {
#pragma unroll
            for (int s = 0; s < kStages; ++s) {
                const int it = it0 + s;
                if (it >= nit) break;
                const int item = items[it];
                if (item & (1 << 26)) {                 // first chunk of a class: its buffer must have been drained
                    if (item & (1 << 20)) { mbar_wait(&s_dempty[1], dph1); dph1 ^= 1; }
                    else { mbar_wait(&s_dempty[0], dph0); dph0 ^= 1; }
                }
                if (pipe == 0 && lane == 0 && it == 20) UM_T(13);
                mbar_wait(&s_full[pipe][s], fph);
                if (pipe == 0 && lane == 0 && it == 20) UM_T(14);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned d_tmem = tmem + (item & 0x10007f);
                const unsigned long long a0 = adesc + (unsigned)((s * kStageBytes) >> 4);
                const unsigned long long b0 = bdesc + (unsigned)((s * kStageBytes) >> 4);
                if (elect_one()) {
                    if (!(dbg & 1)) {
                        umma_bf16(d_tmem, a0, b0, (item & (1 << 24)) ? 0u : 1u);
                        umma_bf16(d_tmem, a0, b0 + (256 >> 4), 1u);
                        umma_bf16(d_tmem, a0 + (2048 >> 4), b0 + (512 >> 4), 1u);
                        umma_bf16(d_tmem, a0 + (2048 >> 4), b0 + (768 >> 4), 1u);
                    }
                    umma_commit(&s_empty[pipe][s]);
                    if (item & (1 << 25)) umma_commit(&s_dfull[(item >> 20) & 1]);
                }
                __syncwarp();
                if (pipe == 0 && lane == 0 && it == 20) UM_T(15);
                if (pipe == 0 && lane == 0 && it == 21) UM_T(19);
                if (pipe == 0 && lane == 0 && it == 0) UM_T(8);
                if (pipe == 0 && lane == 0 && it == 25) UM_T(9);
            }
            fph ^= 1;
        }
        if (pipe == 0 && lane == 0) UM_T(10);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) UM_T(12);
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(kTmemCols) : "memory");
    }
}

// Reference operand images: [tile of 4 references][chunk][1 KB]:
//   [mma = s2 * 2 + which][kc][nr = ref_in_tile * 2 + uv][8 bf16], ring quad j = 2 s2 + kc of the chunk,
//   which 0: [c_hi x4 | c_hi x4], which 1: [c_lo x4 | 0 x4], c = U = re - im (uv 0) or V = re + im (uv 1)
__global__ void um_pack_refs_kernel(const unsigned char* __restrict__ refspec, int R, size_t row_bytes, int nch,
                                    unsigned char* __restrict__ img, int ntile)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;       // (tile, chunk, j, nr)
    const long total = (long)ntile * nch * 4 * 8;
    if (idx >= total) return;
    const int nr = (int)(idx & 7), j = (int)((idx >> 3) & 3);
    const long tc = idx >> 5;
    const int gc = (int)(tc % nch), tile = (int)(tc / nch);
    const int ref = 4 * tile + (nr >> 1), uv = nr & 1;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    if (ref < R) {
        // reference layout (cra_common.cuh): quad j = [re unit | im unit], a unit = bf16 hi x4 then bf16 lo x4
        const __nv_bfloat16* u = reinterpret_cast<const __nv_bfloat16*>(refspec + (size_t)ref * row_bytes + (size_t)gc * 128 + j * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float re = __bfloat162float(u[i]) + __bfloat162float(u[4 + i]);
            const float im = __bfloat162float(u[8 + i]) + __bfloat162float(u[12 + i]);
            c[i] = uv ? re + im : re - im;
        }
    }
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { hi[i] = __float2bfloat16_rn(c[i]); lo[i] = __float2bfloat16_rn(c[i] - __bfloat162float(hi[i])); }
    const int s2 = j >> 1, kc = j & 1;
    unsigned char* o = img + (size_t)tc * kBBytes;
    __nv_bfloat16* o0 = reinterpret_cast<__nv_bfloat16*>(o + ((s2 * 2 + 0) * 2 + kc) * 128 + nr * 16);
    __nv_bfloat16* o1 = reinterpret_cast<__nv_bfloat16*>(o + ((s2 * 2 + 1) * 2 + kc) * 128 + nr * 16);
    const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) { o0[i] = hi[i]; o0[4 + i] = hi[i]; o1[i] = lo[i]; o1[4 + i] = z; }
}

struct Sched { int nring = -1, maxrin = -1, dev = -1, istride = 0, lo = 0, hi = 0; std::vector<int> len; };
Sched g_sched;

// class order: position parity = consumer group; the two single-residue classes both go to group 0.
// Within a class the frequency slots are dealt to the pipelines (heaviest first, least loaded pipeline).
int bind_schedule(const CraRingTab& h, const std::vector<int>& koff, cudaStream_t st)
{
    int dev = 0; cudaGetDevice(&dev);
    std::vector<int> len(h.len, h.len + h.nring);
    if (g_sched.nring == h.nring && g_sched.maxrin == h.maxrin && g_sched.dev == dev && g_sched.len == len) return 0;
    const int cls[9] = {0, 1, 8, 2, 3, 4, 5, 6, 7};
    std::vector<int> items[kPipes];
    long load[kPipes] = {0};
    int lo = 0, hi = 0;
    for (int pos = 0; pos < 9; ++pos) {
        const int n2 = cls[pos];
        if (pos < 8) lo |= n2 << (4 * pos); else hi = n2;
        std::vector<int> ks;                       // frequencies of the class in slot order
        if (n2 == 0) for (int j = 0; j <= 8; ++j) ks.push_back(16 * j);
        else if (n2 == 8) for (int j = 0; j < 8; ++j) ks.push_back(8 + 16 * j);
        else { for (int j = 0; j < 8; ++j) ks.push_back(n2 + 16 * j); for (int j = 0; j < 8; ++j) ks.push_back(16 - n2 + 16 * j); }
        std::vector<int> order(ks.size());
        for (size_t i = 0; i < ks.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return koff[ks[a] + 1] - koff[ks[a]] > koff[ks[b] + 1] - koff[ks[b]]; });
        size_t first_item[kPipes];
        for (int p = 0; p < kPipes; ++p) first_item[p] = items[p].size();
        for (int s : order) {
            const int k = ks[s], c0 = koff[k], c1 = koff[k + 1];
            if (c1 <= c0) { cra_set_error("UMMA CCF kernel: a frequency without rings"); return 1; }
            int p = 0;
            for (int pp = 1; pp < kPipes; ++pp) if (load[pp] < load[p]) p = pp;
            for (int gc = c0; gc < c1; ++gc)
                items[p].push_back((s * 8) | (gc << 7) | ((pos & 1) << 20) | ((gc == c0) ? (1 << 24) : 0) | (1 << 27));
            load[p] += c1 - c0;
        }
        for (int p = 0; p < kPipes; ++p) {
            if (items[p].size() == first_item[p]) { cra_set_error("UMMA CCF kernel: a pipeline without work in a class"); return 1; }
            items[p][first_item[p]] |= 1 << 26;
            items[p].back() |= 1 << 25;
        }
    }
    int istride = 0;
    for (int p = 0; p < kPipes; ++p) istride = std::max(istride, (int)items[p].size());
    if (istride > kMaxItems || koff[N / 2 + 1] > 8191) { cra_set_error("ring table too large for the UMMA CCF schedule"); return 1; }
    static int h_items[kPipes][kMaxItems]; int h_n[kPipes];
    memset(h_items, 0, sizeof(h_items));
    for (int p = 0; p < kPipes; ++p) { h_n[p] = (int)items[p].size(); for (size_t i = 0; i < items[p].size(); ++i) h_items[p][i] = items[p][i]; }
    CRA_CUDA(cudaStreamSynchronize(st));
    CRA_CUDA(cudaMemcpyToSymbol(g_items, h_items, sizeof(h_items)));
    CRA_CUDA(cudaMemcpyToSymbol(g_nitems, h_n, sizeof(h_n)));
    // the copies come from pageable memory on the legacy stream and the kernel runs on a non-blocking stream:
    // make sure they have landed (diagnostic path; one-off per geometry and device)
    CRA_CUDA(cudaDeviceSynchronize());
    g_sched.nring = h.nring; g_sched.maxrin = h.maxrin; g_sched.dev = dev; g_sched.len = len;
    g_sched.istride = istride; g_sched.lo = lo; g_sched.hi = hi;
    return 0;
}

}  // namespace

bool cra_ccf_um_supported(int log2n) { return log2n == LOG2N; }
int cra_ccf_um_num_tiles(int R) { return (R + 3) / 4; }
size_t cra_ccf_um_refimg_bytes(int max_refs, int nch) { return (size_t)((max_refs + 3) / 4) * nch * kBBytes; }

int cra_ccf_um_pack_refs(const unsigned char* refspec, int R, const CraFragTab& frag, unsigned char* img, cudaStream_t st)
{
    const int ntile = (R + 3) / 4;
    const long total = (long)ntile * frag.nch * 32;
    um_pack_refs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(refspec, R, cra_frag_row_bytes(frag.nch), frag.nch, img, ntile);
    CRA_CUDA(cudaGetLastError());
    return 0;
}

int cra_launch_ccf_um(const unsigned char* spec, int nrows, const unsigned char* refimg, int R, const CraRingTab& htab,
                      const CraFragTab& frag, const std::vector<int>& h_koff, const float2* twid, CraCand* cand,
                      int ntile_n, const float2* norm, const float* tref, cudaStream_t st)
{
    if (htab.log2n != LOG2N) { cra_set_error("UMMA CCF kernel: maxrin must be 256"); return 1; }
    if (bind_schedule(htab, h_koff, st)) return 1;
    const size_t smem = (size_t)kPipes * kStages * kStageBytes + (size_t)(N2 - kYT) * N1 * 128 * sizeof(float2) + N * sizeof(float2) +
                        (size_t)kPipes * g_sched.istride * sizeof(int);
    if (cra_ensure_dyn_smem(reinterpret_cast<const void*>(&ccf_um_kernel), smem)) return 1;
    const long ncta_m = (nrows + 31) / 32;
    const long nblk = ncta_m * ntile_n;
    if (nblk <= 0) return 0;
    if (nblk > 2147483647L) { cra_set_error("ccf grid too large; lower row_batch"); return 1; }
    ccf_um_kernel<<<(unsigned)nblk, kThreads, smem, st>>>(spec, nrows, refimg, R, cra_frag_row_bytes(frag.nch), frag.nch, twid, cand,
                                                          ntile_n, norm, tref, g_sched.istride, g_sched.lo, g_sched.hi,
                                                          getenv("CRA_UM_DBG") ? atoi(getenv("CRA_UM_DBG")) : 0);
    CRA_CUDA(cudaGetLastError());
    if (getenv("CRA_UM_DBG") && (atoi(getenv("CRA_UM_DBG")) & 32)) {
        long long h[32];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_dbg, sizeof(h));
        fprintf(stderr, "um clocks:");
        for (int i = 1; i < 21; ++i) fprintf(stderr, " t%d=%lld", i, h[i] ? h[i] - h[0] : -1);
        fprintf(stderr, "\n");
    }
    return 0;
}
