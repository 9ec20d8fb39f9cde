// cra_ccf_mma.cu -- Crosrng_ms ring contraction on the tensor cores, fused with the inverse FFT
// and the peak search (EMAN2 Util::Crosrng_ms + the best-of loop of Util::multiref_polar_ali_2d;
// reference call site test_mref.py:200-201; replaces cu_ccf_mult_m + cuFFT C2R + cu_max_idx_batch,
// cuda/gpu_aln_noref.cu:1009-1143, :2198-2206, :1305-1346).
//
// For every angular frequency k the contraction over rings is a small real GEMM
//     [ (row, re) ; (row, im) ]  x  [ (ref, re) , (ref, im) ]        K = rings reaching k  (<= nring)
// whose four products per (row, ref) are the ring sums A = sum c.x d.x, D = sum c.y d.x,
// C = sum c.x d.y, B = sum c.y d.y of Crosrng_ms (c = weighted reference spectrum, d = particle
// spectrum):  q_k = (A+B) + i(D-C),  t_k = (A-B) - i(C+D).  maxrin/2+1 independent GEMMs with
// K <= 36 are far too small for tcgen05 (M >= 64 per instruction and the 2 KB of q/t spectrum per
// pair must stay on chip for the inverse FFT, which caps a tile at ~100 pairs per SM), so they run as
// warp-level mma.sync.m16n8k16 (bf16, fp32 accumulate) with operands straight from L2 in
// fragment order (cra_common.cuh, CRA_FMT_FRAG): one 256-bit load is a thread's A fragment (hi
// registers then lo registers, already in mma operand order), one 128-bit load its B fragment.  FP32 accuracy comes from the split a = a_hi + a_lo:
// a.b ~= a_hi b_hi + a_hi b_lo + a_lo b_hi (3 MMAs; the dropped a_lo b_lo term is 2^-18 relative).
//
// One CTA = 8 rows (particle x shift) x up to 4*NJ references; 16 warps share the frequencies
// (host-balanced lists in constant memory), each warp keeps one frequency's 8 x 4*NJ tile in
// registers, and writes the Hermitian-extended W = q + i t to shared memory when the last ring
// chunk of that frequency is done.  Then the whole CTA runs one length-maxrin complex inverse FFT
// per pair (two register passes), the ">=" argmax, the straight/mirror choice and the
// best-over-references rule, and stores one (value, code) candidate per row.
#include "cra_common.cuh"
#include "cra_fft.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>

namespace {

constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int kMaxItems = 320;
constexpr int kPfDist = 40;             // row blocks between an L2 prefetch and its use (~148 CTAs in flight / tiles per block)          // chunk items per warp (constant memory)

// per-warp work lists: gc | last << 23   (gc = chunk index within a row); the frequencies a warp
// completes, in order, as packed shared-memory positions of W[k] and W[N-k]: i0 | i1 << 12 | (k != 0, N/2) << 24
constexpr int kMaxFreq = 64;
// The lists live in global memory and are staged in shared memory by every CTA: a register-indexed
// constant load (LDC) per chunk stalled the warps for ~40 % of the contraction (ncu, round 1).
__device__ int g_items[kWarps][kMaxItems];
__device__ int g_nitems[kWarps];
__device__ int g_flush[kWarps][kMaxFreq];

using crafft::fft_reg;

template <int LOG2N>
struct MShape {
    static constexpr int N = 1 << LOG2N;
    static constexpr int L1 = LOG2N / 2;
    static constexpr int L2 = LOG2N - L1;
    static constexpr int N1 = 1 << L1;
    static constexpr int N2 = 1 << L2;
    static constexpr int PS = N1 * (N2 + 1) + ((N1 * (N2 + 1)) % 2 == 0 ? 1 : 0);   // odd float2 stride of one pair
    // rows and reference quads per CTA, bounded by shared memory for W (PS * 8 bytes per pair)
    static constexpr int ROWS = (LOG2N >= 10) ? 4 : 8;
    static constexpr int NJ = (LOG2N <= 8) ? 3 : 1;
    static constexpr int RS = 4 * NJ;                   // pair slots per row
    static constexpr int NP = ROWS * RS;
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

template <int NJ>
struct Operands { Frag8 a; uint4 b[NJ]; };

template <int LOG2N>
__global__ void __launch_bounds__(kThreads, 1)
ccf_mma_kernel(const unsigned char* __restrict__ spec, int nrows, const unsigned char* __restrict__ refspec, int R,
               size_t row_bytes, const float2* __restrict__ twid, CraCand* __restrict__ cand,
               int nquad, int ncta_n, const float2* __restrict__ norm, const float* __restrict__ tref, int istride, int fstride)
{
    using S = MShape<LOG2N>;
    constexpr int N = S::N, N1 = S::N1, N2 = S::N2, PS = S::PS, NJ = S::NJ, RS = S::RS, ROWS = S::ROWS, NP = S::NP;
    extern __shared__ __align__(16) float2 s_dyn[];
    float2* s_w = s_dyn;                      // NP * PS
    float2* s_tw = s_dyn + NP * PS;           // N : s_tw[j*N2 + n2] = exp(+2 pi i n2 j / N)
    int* s_items = reinterpret_cast<int*>(s_tw + N);            // kWarps * istride chunk items
    int* s_flush = s_items + kWarps * istride;                  // kWarps * fstride completed frequencies
    __shared__ CraCand s_pair[NP];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    // reference tile fastest so that concurrently resident CTAs share the row spectra in L2
    const int cn = blockIdx.x % ncta_n, cm = blockIdx.x / ncta_n;
    const int qbase = nquad / ncta_n, qrem = nquad % ncta_n;
    const int nj = qbase + (cn < qrem ? 1 : 0);                 // reference quads of this CTA (<= NJ)
    const int q0 = cn * qbase + min(cn, qrem);
    const int row0 = cm * ROWS;
    for (int i = tid; i < N; i += kThreads) s_tw[i] = twid[i];
    {
        const int nit = g_nitems[warp];
        for (int i = lane; i < nit; i += 32) s_items[warp * istride + i] = g_items[warp][i];
        for (int i = lane; i < fstride; i += 32) s_flush[warp * fstride + i] = g_flush[warp][i];
        __syncwarp();
    }

    // The row spectra were written by the row kernel a whole batch ago and sit in HBM: the first
    // reference tile of a row block pulls the block that will be needed kPfDist blocks later into L2.
    if (cn == 0) {
        const long r0 = (long)(cm + kPfDist) * ROWS;
        if (r0 < nrows) {
            const long r1 = min((long)nrows, r0 + ROWS);
            const unsigned char* p0 = spec + (size_t)r0 * row_bytes;
            const size_t nline = (size_t)(r1 - r0) * row_bytes / 128;
            for (size_t i = tid; i < nline; i += kThreads)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(p0 + i * 128));
        }
    }
    // ---- contraction: this warp's frequencies, chunk by chunk ---------------------------------
    // All NJ quads are always multiplied (a tile with fewer live quads re-reads its quad 0; the
    // dead pair slots are skipped by the FFT passes below): no predication around the MMAs.
    {
        int rrow = row0 + (ROWS == 8 ? g : (g & 3));
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
        const bool wr = (ROWS == 8) || (g < 4);
        float2* const wbase = s_w + (g * RS + t) * PS;

        // a.b ~= a_hi b_lo + a_lo b_hi + a_hi b_hi, small terms first, one FP32 accumulator
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
            const int i0 = f & 4095, i1 = (f >> 12) & 4095;
            const bool two = (f >> 24) != 0;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                // c0 = A (re.re), c1 = D (row re . ref im), c2 = C (row im . ref re), c3 = B (im.im)
                const float A = acc[j][0], D = acc[j][1], C = acc[j][2], B = acc[j][3];
                // s = (A+B, A-B), tv = (C+D, D-C);  W[k] = s + tv,  W[N-k] = s - tv
                const float sx = A + B, sy = A - B, tx = C + D, ty = D - C;
                if (wr) {
                    wbase[4 * j * PS + i0] = make_float2(sx + tx, sy + ty);
                    if (two) wbase[4 * j * PS + i1] = make_float2(sx - tx, sy - ty);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
            }
        };

        // three operand sets rotate: two chunk loads are always in flight behind the one being multiplied
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
#undef CRA_LOAD_OPS
#undef CRA_COMPUTE
    }
    __syncthreads();

    // ---- inverse FFT of every pair, pass 1: (pair, n2): N1-point DFT over n1, twiddle ----------
    const int npr = 4 * nj;                   // live pair slots per row
    const int nlive = ROWS * npr;
    for (int item = tid; item < nlive * N2; item += kThreads) {
        const int pi = item / N2, n2 = item - pi * N2;
        const int pr = pi / npr, pc = pi - pr * npr;
        float2* w = s_w + (pr * RS + pc) * PS + n2;
        float2 x[N1];
#pragma unroll
        for (int j = 0; j < N1; ++j) x[j] = w[j * (N2 + 1)];
        fft_reg<N1, 1>(x);
#pragma unroll
        for (int j = 0; j < N1; ++j) {
            if (j == 0) { w[0] = x[0]; continue; }
            const float2 tw = s_tw[j * N2 + n2];
            w[j * (N2 + 1)] = crafft::cmul(x[j], tw);
        }
    }
    __syncthreads();

    // ---- pass 2: (pair, k1): N2-point DFT over n2 -> X[k1 + N1*k2]; argmax over lags ------------
    for (int item = tid; item < nlive * N1; item += kThreads) {
        const int pi = item / N1, k1 = item - pi * N1;
        const int pr = pi / npr, pc = pi - pr * npr;
        const int slot = pr * RS + pc;
        const float2* w = s_w + slot * PS + k1 * (N2 + 1);
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
        // the N1 lanes of one pair are consecutive and aligned inside a warp (N1 <= 32)
#pragma unroll
        for (int o = N1 >> 1; o > 0; o >>= 1) {
            float oq = __shfl_xor_sync(0xffffffffu, bq, o); int omq = __shfl_xor_sync(0xffffffffu, mq, o);
            float ot = __shfl_xor_sync(0xffffffffu, bt, o); int omt = __shfl_xor_sync(0xffffffffu, mt, o);
            if (better(oq, omq, bq, mq)) { bq = oq; mq = omq; }
            if (better(ot, omt, bt, mt)) { bt = ot; mt = omt; }
        }
        if (k1 == 0) {
            const int row = row0 + pr, ref = 4 * q0 + pc;
            CraCand cd;
            if (row < nrows && ref < R) {
                // deferred Normalize_ring (cra_common.cuh): the row spectrum is stored un-normalised;
                // (x - avg)/sigma only moves the DC bins, i.e. shifts every lag by -avg * tref[ref]
                const float2 nm = norm[row];
                const float sc = nm.y / (float)N, dc = nm.x * tref[ref];
                const float qn = (bq - dc) * sc, qm = (bt - dc) * sc;
                if (qn >= qm) { cd.v = qn; cd.code = ref * 8192 + (mq + 1); }
                else          { cd.v = qm; cd.code = ref * 8192 + 4096 + (mt + 1); }
            } else { cd.v = -INFINITY; cd.code = -1; }
            s_pair[slot] = cd;
        }
    }
    __syncthreads();
    if (tid < ROWS) {
        const int row = row0 + tid;
        if (row < nrows) {
            CraCand best; best.v = -INFINITY; best.code = -1;
            for (int n = 0; n < npr; ++n) {
                const CraCand c = s_pair[tid * RS + n];
                if (c.code >= 0 && c.v >= best.v) best = c;
            }
            cand[(size_t)row * ncta_n + cn] = best;
        }
    }
}

struct Sched { std::vector<int> koff; int nring = -1, maxrin = -1, dev = -1, istride = 0, fstride = 0; std::vector<int> len; };
Sched g_sched;

// Balance the frequencies over the warps (longest-processing-time first) and upload the lists.
int bind_schedule(const CraRingTab& h, const std::vector<int>& koff, cudaStream_t st)
{
    int dev = 0; cudaGetDevice(&dev);
    std::vector<int> len(h.len, h.len + h.nring);
    if (g_sched.nring == h.nring && g_sched.maxrin == h.maxrin && g_sched.dev == dev && g_sched.len == len) return 0;
    const int nk = h.maxrin / 2 + 1;
    std::vector<std::vector<int>> lists(kWarps), freqs(kWarps);
    std::vector<int> load(kWarps, 0);
    int maxc = 0;
    for (int k = 0; k < nk; ++k) maxc = std::max(maxc, koff[k + 1] - koff[k]);
    for (int c = maxc; c >= 1; --c)
        for (int k = 0; k < nk; ++k) {
            if (koff[k + 1] - koff[k] != c) continue;
            int w = 0;
            for (int i = 1; i < kWarps; ++i) if (load[i] < load[w]) w = i;
            for (int j = 0; j < c; ++j) lists[w].push_back((koff[k] + j) | ((j == c - 1) ? (1 << 23) : 0));
            freqs[w].push_back(k);
            load[w] += c;
        }
    static int h_items[kWarps][kMaxItems]; int h_n[kWarps];
    static int h_flush[kWarps][kMaxFreq];
    memset(h_items, 0, sizeof(h_items)); memset(h_flush, 0, sizeof(h_flush));
    const int L2 = h.log2n - h.log2n / 2, N2 = 1 << L2, N = h.maxrin;       // MShape: k -> (k >> L2) * (N2 + 1) + (k & (N2 - 1))
    for (int w = 0; w < kWarps; ++w) {
        if ((int)lists[w].size() > kMaxItems || (int)freqs[w].size() > kMaxFreq || koff[nk] > 8191 || nk > 1024) { cra_set_error("ring table too large for the tensor-core CCF schedule"); return 1; }
        h_n[w] = (int)lists[w].size();
        for (size_t i = 0; i < lists[w].size(); ++i) h_items[w][i] = lists[w][i];
        for (size_t i = 0; i < freqs[w].size(); ++i) {
            const int k = freqs[w][i], kk = (N - k) & (N - 1);
            const int i0 = (k >> L2) * (N2 + 1) + (k & (N2 - 1)), i1 = (kk >> L2) * (N2 + 1) + (kk & (N2 - 1));
            h_flush[w][i] = i0 | (i1 << 12) | ((k != 0 && k != N / 2) ? (1 << 24) : 0);
        }
    }
    CRA_CUDA(cudaStreamSynchronize(st));
    CRA_CUDA(cudaMemcpyToSymbol(g_items, h_items, sizeof(h_items)));
    CRA_CUDA(cudaMemcpyToSymbol(g_nitems, h_n, sizeof(h_n)));
    CRA_CUDA(cudaMemcpyToSymbol(g_flush, h_flush, sizeof(h_flush)));
    // the copies come from pageable memory on the legacy stream and the kernel runs on a non-blocking stream:
    // make sure they have landed (diagnostic path; one-off per geometry and device)
    CRA_CUDA(cudaDeviceSynchronize());
    g_sched.istride = 0; g_sched.fstride = 0;
    for (int w = 0; w < kWarps; ++w) {
        g_sched.istride = std::max(g_sched.istride, (int)lists[w].size());
        g_sched.fstride = std::max(g_sched.fstride, (int)freqs[w].size());
    }
    g_sched.nring = h.nring; g_sched.maxrin = h.maxrin; g_sched.dev = dev; g_sched.len = len;
    return 0;
}

template <int LOG2N>
int launch_t(const unsigned char* spec, int nrows, const unsigned char* refspec, int R, size_t row_bytes,
             const float2* twid, CraCand* cand, int ntile_n, const float2* norm, const float* tref, cudaStream_t st)
{
    using S = MShape<LOG2N>;
    const size_t smem = ((size_t)S::NP * S::PS + S::N) * sizeof(float2) + (size_t)kWarps * (g_sched.istride + g_sched.fstride) * sizeof(int);
    if (cra_ensure_dyn_smem(reinterpret_cast<const void*>(&ccf_mma_kernel<LOG2N>), smem)) return 1;
    const int nquad = (R + 3) / 4;
    const long ncta_m = (nrows + S::ROWS - 1) / S::ROWS;
    const long nblk = ncta_m * ntile_n;
    if (nblk <= 0) return 0;
    if (nblk > 2147483647L) { cra_set_error("ccf grid too large; lower row_batch"); return 1; }
    ccf_mma_kernel<LOG2N><<<(unsigned)nblk, kThreads, smem, st>>>(spec, nrows, refspec, R, row_bytes, twid, cand, nquad, ntile_n, norm, tref, g_sched.istride, g_sched.fstride);
    CRA_CUDA(cudaGetLastError());
    return 0;
}

template <int LOG2N> int num_tiles_t(int R) { const int nq = (R + 3) / 4; return (nq + MShape<LOG2N>::NJ - 1) / MShape<LOG2N>::NJ; }

}  // namespace

// reference tiles (candidates per row) the kernel produces for R references
int cra_ccf_mma_num_tiles(int R, int log2n)
{
    switch (log2n) {
        case 5: return num_tiles_t<5>(R);   case 6: return num_tiles_t<6>(R);   case 7: return num_tiles_t<7>(R);
        case 8: return num_tiles_t<8>(R);   case 9: return num_tiles_t<9>(R);   default: return num_tiles_t<10>(R);
    }
}

int cra_launch_ccf_mma(const unsigned char* spec, int nrows, const unsigned char* refspec, int R, const CraRingTab& htab,
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
        case 10: return launch_t<10>(spec, nrows, refspec, R, rb, twid, cand, ntile_n, norm, tref, st);
        default: cra_set_error("maxrin must be a power of two in [32, 1024]"); return 1;
    }
}
