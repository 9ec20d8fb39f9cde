// cra_ccf_rr.cu -- "reference-resident" variant of the Crosrng_ms + inverse FFT + peak kernel
// (same arithmetic and output as ccf_peak_kernel in cra_ccf.cu; EMAN2 Util::Crosrng_ms, reference
// call site test_mref.py:200-201).
//
// Why: ncu shows ccf_peak_kernel pulling 7.2 TB/s through L2 -> SM during the contraction (each
// 16-pair CTA streams 94 KB of row spectra and 94 KB of reference spectra), i.e. it is bound by that
// bandwidth and by the bytes it can keep in flight.  Here a persistent CTA keeps ONE reference group
// (4 references, 2*nc float4 = 94 KB at ou=36) resident in shared memory and streams row groups
// past it, so the L2 -> SM traffic per pair halves (5.9 KB).  Three 96-thread sub-tiles per CTA work
// on three different row groups against the same resident references, synchronising only among
// themselves (named barriers).  Row spectra are prefetched RD rings ahead through registers.
#include "cra_common.cuh"
#include "cra_fft.cuh"
#include <math.h>
#include <string.h>

namespace {

using crafft::fft_reg;

constexpr int TM = 4, TN = 4, NP = TM * TN;
constexpr int NSUBR = 3;     // sub-tiles (row groups in flight) per CTA
constexpr int RD = 6;        // row-spectrum prefetch depth (rings)

__constant__ int r_coff[CRA_MAX_RINGS];
__constant__ int r_half[CRA_MAX_RINGS];

template <int LOG2N>
struct RShape {
    static constexpr int N = 1 << LOG2N;
    static constexpr int L1 = LOG2N / 2;
    static constexpr int L2 = LOG2N - L1;
    static constexpr int N1 = 1 << L1;
    static constexpr int N2 = 1 << L2;
    static constexpr int NQ = N / 4;
    static constexpr int NW = 3 * N / 8;
    static constexpr int NT1 = ((NW + 31) / 32) * 32 < 32 ? 32 : ((NW + 31) / 32) * 32;
    static constexpr int NT = NT1 * NSUBR;
    static constexpr int PS = N1 * (N2 + 1);
};

__device__ __forceinline__ bool better(float v, int m, float bv, int bm)
{
    return (v > bv) || (v == bv && m > bm);
}
__device__ __forceinline__ void sub_barrier(int sub, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(sub + 1), "r"(nthreads) : "memory");
}

struct Acc { float A[TM][TN], B[TM][TN], C[TM][TN], D[TM][TN]; };

__device__ __forceinline__ void ring_fma(Acc& a, const float4 (&d)[2], const float4 (&c)[2])
{
    const float dx[TM] = {d[0].x, d[0].z, d[1].x, d[1].z}, dy[TM] = {d[0].y, d[0].w, d[1].y, d[1].w};
    const float cx[TN] = {c[0].x, c[0].z, c[1].x, c[1].z}, cy[TN] = {c[0].y, c[0].w, c[1].y, c[1].w};
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
        for (int n = 0; n < TN; ++n) {
            a.A[m][n] = fmaf(cx[n], dx[m], a.A[m][n]);
            a.B[m][n] = fmaf(cy[n], dy[m], a.B[m][n]);
            a.C[m][n] = fmaf(cx[n], dy[m], a.C[m][n]);
            a.D[m][n] = fmaf(cy[n], dx[m], a.D[m][n]);
        }
}

// ring sums of 16 pairs at frequency k (complex elements k < len/2 only), then W[k], W[N-k]
template <int LOG2N>
__device__ __forceinline__ void contract_freq(int k, int nring, const float4* __restrict__ dq,
                                              const float4* __restrict__ cs, float2* __restrict__ s_w)
{
    using S = RShape<LOG2N>;
    constexpr int N = S::N, N2 = S::N2, PS = S::PS;
    Acc a;
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
        for (int n = 0; n < TN; ++n) { a.A[m][n] = 0.f; a.B[m][n] = 0.f; a.C[m][n] = 0.f; a.D[m][n] = 0.f; }

    float4 dbuf[RD][2];
    bool valid[RD];
    int i = nring - 1;
#pragma unroll
    for (int s = 0; s < RD; ++s) {
        const int j = i - s, jj = j < 0 ? 0 : j;
        valid[s] = (j >= 0) && (k < r_half[jj]);
        if (valid[s]) {
            const int e = 2 * r_coff[jj] + k;
            dbuf[s][0] = __ldg(dq + e); dbuf[s][1] = __ldg(dq + e + r_half[jj] + 1);
        }
    }
    bool more = true;
    while (more) {
#pragma unroll
        for (int s = 0; s < RD; ++s) {
            if (more && valid[s]) {
                const int j = i - s;
                const int e = 2 * r_coff[j] + k, p = r_half[j] + 1;
                float4 c[2];
                c[0] = cs[e]; c[1] = cs[e + p];
                ring_fma(a, dbuf[s], c);
                const int jn = j - RD, jj = jn < 0 ? 0 : jn;
                const bool vn = (jn >= 0) && (k < r_half[jj]);
                if (vn) {
                    const int en = 2 * r_coff[jj] + k;
                    dbuf[s][0] = __ldg(dq + en); dbuf[s][1] = __ldg(dq + en + r_half[jj] + 1);
                }
                valid[s] = vn;
            } else {
                more = false;       // ring lengths only shrink: the first invalid slot ends the walk
            }
        }
        i -= RD;
    }
    const int kk = (N - k) & (N - 1);
    const int i0 = (k >> S::L2) * (N2 + 1) + (k & (N2 - 1));
    const int i1 = (kk >> S::L2) * (N2 + 1) + (kk & (N2 - 1));
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
        for (int n = 0; n < TN; ++n) {
            float2* w = s_w + (m * TN + n) * PS;
            const float A = a.A[m][n], B = a.B[m][n], C = a.C[m][n], D = a.D[m][n];
            w[i0] = make_float2(A + B + C + D, A - B + D - C);
            if (k != 0) w[i1] = make_float2(A + B - C - D, A - B + C - D);
        }
}

template <int LOG2N>
__global__ void __maxnreg__(216)
ccf_rr_kernel(const float4* __restrict__ spec, int nrows, const float4* __restrict__ refspec, int R,
              int nring, int nc, const float2* __restrict__ twid, CraCand* __restrict__ cand, int ntile_n,
              int ngroups, long units_per_cta)
{
    using S = RShape<LOG2N>;
    constexpr int N = S::N, N1 = S::N1, N2 = S::N2, NT = S::NT, NT1 = S::NT1, PS = S::PS, NQ = S::NQ, NW = S::NW;
    extern __shared__ __align__(16) unsigned char rr_smem[];
    float4* s_ref = reinterpret_cast<float4*>(rr_smem);                        // 2*nc float4: the resident reference group
    float2* s_wall = reinterpret_cast<float2*>(s_ref + 2 * (size_t)nc);        // NSUBR * NP * PS
    float2* s_tw = s_wall + (size_t)NSUBR * NP * PS;                           // N
    __shared__ CraCand s_pair_all[NSUBR][NP];
    __shared__ float s_nyq_all[NSUBR][NP][8];

    const int sub = threadIdx.x / NT1, tid = threadIdx.x - sub * NT1;
    float2* s_w = s_wall + (size_t)sub * NP * PS;
    CraCand* s_pair = s_pair_all[sub];
    float (*s_nyq)[8] = s_nyq_all[sub];
    for (int i = threadIdx.x; i < N; i += NT) s_tw[i] = twid[i];

    const long total = (long)ntile_n * ngroups;
    long u = (long)blockIdx.x * units_per_cta;
    const long u1 = (u + units_per_cta < total) ? u + units_per_cta : total;
    while (u < u1) {                                   // CTA-uniform: one segment = one reference group
        const int tn = (int)(u / ngroups), g0 = (int)(u - (long)tn * ngroups);
        const int gend = (int)(((long)ngroups - g0 < u1 - u) ? ngroups : g0 + (u1 - u));
        __syncthreads();                               // previous segment's readers of s_ref are done
        const float4* cq = refspec + (size_t)tn * nc * 2;
        for (int i = threadIdx.x; i < 2 * nc; i += NT) s_ref[i] = __ldg(cq + i);
        __syncthreads();

        for (int g = g0 + sub; g < gend; g += NSUBR) {
            const float4* dq = spec + (size_t)g * nc * 2;
            if (tid < NW) {
                contract_freq<LOG2N>(tid, nring, dq, s_ref, s_w);
                if (tid >= NQ) contract_freq<LOG2N>(tid + N / 8, nring, dq, s_ref, s_w);
            }
            // real-valued ring elements: frequency N/2 of full rings (last warp of the sub-tile) and
            // the Nyquist term of shorter rings (warp 1 % nwarps), folded into W after the barrier
            const int nl = tid - ((NT1 >= 64) ? 32 : 0);
            const bool nyq_lane = (nl >= 0 && nl < 32);
            {
                const float2* d2 = reinterpret_cast<const float2*>(dq);
                const float2* c2 = reinterpret_cast<const float2*>(s_ref);
                const int l = tid - (NT1 - 32);
                if (l >= 0 && l < NP) {
                    const int m = l / TN, n = l % TN;
                    float a = 0.f;
                    for (int i = nring - 1; i >= 0 && r_half[i] == N / 2; --i)
                        a = fmaf(c2[cra_spec_idx(r_coff[i], N / 2, n, N / 2)].x,
                                 __ldg(d2 + cra_spec_idx(r_coff[i], N / 2, m, N / 2)).x, a);
                    const int h = N / 2;
                    s_w[l * PS + (h >> S::L2) * (N2 + 1) + (h & (N2 - 1))] = make_float2(a, a);
                }
                if (nyq_lane) {
                    const int pair = nl & (NP - 1), part = nl / NP, m = pair / TN, n = pair % TN;
                    int cls = -1, cur = -1; float a = 0.f;
                    for (int i = 0; i < nring && r_half[i] < N / 2; ++i) {
                        const int h = r_half[i];
                        if (h != cur) {
                            if (cls >= 0 && (cls & 1) == part) s_nyq[pair][cls & 7] = a;
                            cur = h; ++cls; a = 0.f;
                        }
                        if ((cls & 1) == part)
                            a = fmaf(c2[cra_spec_idx(r_coff[i], h, n, h)].x, __ldg(d2 + cra_spec_idx(r_coff[i], h, m, h)).x, a);
                    }
                    if (cls >= 0 && (cls & 1) == part) s_nyq[pair][cls & 7] = a;
                }
            }
            sub_barrier(sub, NT1);
            if (nyq_lane) {
                const int pair = nl & (NP - 1), part = nl / NP;
                float2* w = s_w + pair * PS;
                int cls = -1, cur = -1;
                for (int i = 0; i < nring && r_half[i] < N / 2; ++i) {
                    const int h = r_half[i];
                    if (h == cur) continue;
                    cur = h; ++cls;
                    if ((cls & 1) != part) continue;
                    const float a = s_nyq[pair][cls & 7];
                    const int hh = N - h;
                    float2* p0 = w + (h >> S::L2) * (N2 + 1) + (h & (N2 - 1));
                    float2* p1 = w + (hh >> S::L2) * (N2 + 1) + (hh & (N2 - 1));
                    float2 v = *p0; v.x += a; v.y += a; *p0 = v;
                    v = *p1; v.x += a; v.y += a; *p1 = v;
                }
            }
            sub_barrier(sub, NT1);
            // pass 1
            for (int item = tid; item < NP * N2; item += NT1) {
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
                    w[j * (N2 + 1)] = make_float2(x[j].x * t.x - x[j].y * t.y, x[j].x * t.y + x[j].y * t.x);
                }
            }
            sub_barrier(sub, NT1);
            // pass 2 + argmax
            for (int item = tid; item < NP * N1; item += NT1) {
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
                    const int row = g * TM + m, ref = tn * TN + n;
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
            sub_barrier(sub, NT1);
            if (tid < TM) {
                const int row = g * TM + tid;
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
        u += (gend - g0);
    }
}

int bind_rr_tables(const CraRingTab& h, cudaStream_t st)
{
    static int cur_coff[CRA_MAX_RINGS], cur_half[CRA_MAX_RINGS], cur_n = -1, cur_dev = -1;
    int half[CRA_MAX_RINGS];
    for (int i = 0; i < h.nring; ++i) half[i] = h.len[i] >> 1;
    int dev = 0; cudaGetDevice(&dev);
    if (cur_n == h.nring && cur_dev == dev && memcmp(cur_coff, h.coff, sizeof(int) * h.nring) == 0 &&
        memcmp(cur_half, half, sizeof(int) * h.nring) == 0) return 0;
    CRA_CUDA(cudaStreamSynchronize(st));
    CRA_CUDA(cudaMemcpyToSymbol(r_coff, h.coff, sizeof(int) * h.nring));
    CRA_CUDA(cudaMemcpyToSymbol(r_half, half, sizeof(int) * h.nring));
    memcpy(cur_coff, h.coff, sizeof(int) * h.nring); memcpy(cur_half, half, sizeof(int) * h.nring);
    cur_n = h.nring; cur_dev = dev;
    return 0;
}

template <int LOG2N>
int launch_rr(const float* spec, int nrows, const float* refspec, int R, const CraRingTab& h,
              const float2* twid, CraCand* cand, int ntile_n, cudaStream_t st, bool* ran)
{
    using S = RShape<LOG2N>;
    const size_t smem = (size_t)2 * h.nc * sizeof(float4) + ((size_t)NSUBR * NP * S::PS + S::N) * sizeof(float2);
    *ran = false;
    if (smem > 227 * 1024 - 2048) return 0;               // does not fit: caller falls back
    static size_t configured = 0; static int sms = 0;
    if (smem > configured) {
        CRA_CUDA(cudaFuncSetAttribute(ccf_rr_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0; cudaGetDevice(&dev);
        CRA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        configured = smem;
    }
    if (bind_rr_tables(h, st)) return 1;
    const int ngroups = (nrows + TM - 1) / TM;
    const long total = (long)ntile_n * ngroups;
    if (total <= 0) { *ran = true; return 0; }
    const int grid = (int)(total < sms ? total : sms);
    const long units = (total + grid - 1) / grid;
    ccf_rr_kernel<LOG2N><<<grid, S::NT, smem, st>>>(reinterpret_cast<const float4*>(spec), nrows,
                                                   reinterpret_cast<const float4*>(refspec), R, h.nring, h.nc, twid,
                                                   cand, ntile_n, ngroups, units);
    CRA_CUDA(cudaGetLastError());
    *ran = true;
    return 0;
}

}  // namespace

// Returns 0 on success; *ran tells whether this variant applies to the configuration.
int cra_launch_ccf_rr(const float* spec, int nrows, const float* refspec, int R, const CraRingTab& htab,
                      const float2* twid, CraCand* cand, int ntile_n, cudaStream_t st, bool* ran)
{
    *ran = false;
    switch (htab.log2n) {
        case 5: return launch_rr<5>(spec, nrows, refspec, R, htab, twid, cand, ntile_n, st, ran);
        case 6: return launch_rr<6>(spec, nrows, refspec, R, htab, twid, cand, ntile_n, st, ran);
        case 7: return launch_rr<7>(spec, nrows, refspec, R, htab, twid, cand, ntile_n, st, ran);
        case 8: return launch_rr<8>(spec, nrows, refspec, R, htab, twid, cand, ntile_n, st, ran);
        default: return 0;
    }
}
