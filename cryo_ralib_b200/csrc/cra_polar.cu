// cra_polar.cu -- stage 1+2 of the hot path: batched polar resampler
// (Util.Polar2Dm / alrl_ms + quadri), Normalize_ring, per-ring real FFT (Util.Frngs)
// and, for references, Applyws.  Reference call sites: test_mref.py:172-174 (refs),
// :200-201 -> EMAN2 Util::multiref_polar_ali_2d inner loop (particles).
// Replaces cu_resample_to_polar + cuFFT R2C (cuda/gpu_aln_noref.cu:818-879, :1816).
//
// One CTA per (particle, aligned sub-group of RPB consecutive shift rows).  The image tile
// is staged once in shared memory with 128-bit coalesced loads and serves all rows of the
// sub-group; the 6-tap quadratic interpolation gathers from shared memory; the ring sums of
// Normalize_ring are reduced with warp shuffles.  Every ring is then transformed in shared
// memory by two register passes (NA x NB points, radix-2 butterflies with immediate
// twiddles) over flat work lists that keep all threads busy regardless of ring length, and the
// real-FFT split, Normalize_ring's affine map (applied to the spectrum: FFT is linear) or
// the Applyws weights, and the store into the interleaved device spectrum are one final pass
// with 16/32-byte vector stores.
#include "cra_common.cuh"
#include "cra_fft.cuh"
#include <cuda_bf16.h>

namespace {

// split-bf16 of four values: (hi01, hi23, lo01, lo23), value = hi + lo
// A-operand order of a particle row (cra_common.cuh): hi{re01, im01, re23, im23}, lo{same}
__device__ __forceinline__ void split_row_unit(const float (&re)[4], const float (&im)[4], uint4& hi, uint4& lo)
{
    const __nv_bfloat162 hr01 = __floats2bfloat162_rn(re[0], re[1]), hi01 = __floats2bfloat162_rn(im[0], im[1]);
    const __nv_bfloat162 hr23 = __floats2bfloat162_rn(re[2], re[3]), hi23 = __floats2bfloat162_rn(im[2], im[3]);
    const float2 fr01 = __bfloat1622float2(hr01), fi01 = __bfloat1622float2(hi01);
    const float2 fr23 = __bfloat1622float2(hr23), fi23 = __bfloat1622float2(hi23);
    const __nv_bfloat162 lr01 = __floats2bfloat162_rn(re[0] - fr01.x, re[1] - fr01.y), li01 = __floats2bfloat162_rn(im[0] - fi01.x, im[1] - fi01.y);
    const __nv_bfloat162 lr23 = __floats2bfloat162_rn(re[2] - fr23.x, re[3] - fr23.y), li23 = __floats2bfloat162_rn(im[2] - fi23.x, im[3] - fi23.y);
    hi.x = *reinterpret_cast<const unsigned int*>(&hr01); hi.y = *reinterpret_cast<const unsigned int*>(&hi01);
    hi.z = *reinterpret_cast<const unsigned int*>(&hr23); hi.w = *reinterpret_cast<const unsigned int*>(&hi23);
    lo.x = *reinterpret_cast<const unsigned int*>(&lr01); lo.y = *reinterpret_cast<const unsigned int*>(&li01);
    lo.z = *reinterpret_cast<const unsigned int*>(&lr23); lo.w = *reinterpret_cast<const unsigned int*>(&li23);
}

__device__ __forceinline__ uint4 split_bf16x4(float v0, float v1, float v2, float v3)
{
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v0, v1), h23 = __floats2bfloat162_rn(v2, v3);
    const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(v0 - f01.x, v1 - f01.y), l23 = __floats2bfloat162_rn(v2 - f23.x, v3 - f23.y);
    uint4 o;
    o.x = *reinterpret_cast<const unsigned int*>(&h01); o.y = *reinterpret_cast<const unsigned int*>(&h23);
    o.z = *reinterpret_cast<const unsigned int*>(&l01); o.w = *reinterpret_cast<const unsigned int*>(&l23);
    return o;
}

using crafft::cmul;
using crafft::fft_reg;

constexpr int kPolarThreads = 256;

__device__ __forceinline__ float quadri_smem(float x, float y, int nx, const float* __restrict__ f)
{
    // circular closure into [1, nx+1)
    const float fn = (float)nx;
    while (x < 1.0f) x += fn;
    while (x >= fn + 1.0f) x -= fn;
    while (y < 1.0f) y += fn;
    while (y >= fn + 1.0f) y -= fn;
    int i = (int)x, j = (int)y;
    float dx0 = x - i, dy0 = y - j;
    int ip1 = i + 1, im1 = i - 1, jp1 = j + 1, jm1 = j - 1;
    if (ip1 > nx) ip1 -= nx;
    if (im1 < 1) im1 += nx;
    if (jp1 > nx) jp1 -= nx;
    if (jm1 < 1) jm1 += nx;
    const int r0 = (j - 1) * nx - 1;
    float f0 = f[r0 + i];
    float c1 = f[r0 + ip1] - f0;
    float c2 = (c1 - f0 + f[r0 + im1]) * 0.5f;
    float c3 = f[(jp1 - 1) * nx - 1 + i] - f0;
    float c4 = (c3 - f0 + f[(jm1 - 1) * nx - 1 + i]) * 0.5f;
    // dx0, dy0 >= 0 always => hxc = hyc = 1, corner is (i+1, j+1)
    float c5 = f[(jp1 - 1) * nx - 1 + ip1] - f0 - c1 - c3;
    return f0 + dx0 * (c1 + (dx0 - 1.0f) * c2 + dy0 * c5) + dy0 * (c3 + (dy0 - 1.0f) * c4);
}

// same arithmetic, for sample points whose 3x3 neighbourhood is known to lie inside the frame
__device__ __forceinline__ float quadri_inside(float x, float y, int nx, const float* __restrict__ f)
{
    const int i = (int)x, j = (int)y;
    const float dx0 = x - (float)i, dy0 = y - (float)j;
    const float* p = f + (j - 1) * nx + (i - 1);
    const float f0 = p[0];
    const float c1 = p[1] - f0;
    const float c2 = (c1 - f0 + p[-1]) * 0.5f;
    const float c3 = p[nx] - f0;
    const float c4 = (c3 - f0 + p[-nx]) * 0.5f;
    const float c5 = p[nx + 1] - f0 - c1 - c3;
    return f0 + dx0 * (c1 + (dx0 - 1.0f) * c2 + dy0 * c5) + dy0 * (c3 + (dy0 - 1.0f) * c4);
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// pass A of the n = NA*NB point forward FFT of one ring: column b, NA-point DFT, twiddle
template <int NA, int NB>
__device__ __forceinline__ void pass_a(float2* __restrict__ z, int b, float2 base)
{
    float2 x[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) x[a] = z[a * (NB + 1) + b];
    fft_reg<NA, -1>(x);
    z[b] = x[0];
    float2 p = base;
#pragma unroll
    for (int a = 1; a < NA; ++a) {
        z[a * (NB + 1) + b] = cmul(x[a], p);
        if (a + 1 < NA) p = cmul(p, base);
    }
}
// pass B: row ka, NB-point DFT in place; output X[ka + NA*kb] stays at z[ka*(NB+1) + kb]
// (rows of the ring buffer are padded by one complex so that both passes are conflict-free)
template <int NA, int NB>
__device__ __forceinline__ void pass_b(float2* __restrict__ z, int ka)
{
    float2 x[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) x[b] = z[ka * (NB + 1) + b];
    fft_reg<NB, -1>(x);
#pragma unroll
    for (int b = 0; b < NB; ++b) z[ka * (NB + 1) + b] = x[b];
}

struct Chunk { int part, row0, nrow; };

// MODE: 0 = particle rows (optional Normalize_ring), 1 = references (Applyws)
// FMT: CRA_FMT_F32 = planar-pair float2 device spectrum, CRA_FMT_FRAG = split-bf16 fragment layout
// GIMG: the image does not fit into shared memory beside the polar rows (boxes beyond ~170 pixels): the six taps of
// every sample are read from global memory instead (L1 / L2; one particle is read by its few CTAs only), same arithmetic
template <int MODE, int RPB, int FMT, bool GIMG>
__global__ void __launch_bounds__(kPolarThreads)
polar_fft_kernel(const float* __restrict__ images, int nx, const CraRingTab* __restrict__ tab,
                 const float4* __restrict__ samp, const float* __restrict__ sampw,
                 const float2* __restrict__ twid, CraPolarItems items, CraRowMap map,
                 float fix_cx, float fix_cy, int normalize_ring, float* __restrict__ spec, CraFragTab frag,
                 float2* __restrict__ norm, float* __restrict__ tref)
{
    extern __shared__ __align__(16) float smem[];
    const int npix = nx * nx;
    const int lcirc = tab->lcirc;
    const int lcp = tab->lcpad;
    const int maxrin = tab->maxrin;
    float* s_img = smem;                                     // npix (padded to 4); absent with GIMG
    float* s_circ = smem + (GIMG ? 0 : ((npix + 3) & ~3));   // RPB * lcp
    float2* s_tw = reinterpret_cast<float2*>(s_circ + RPB * lcp);   // maxrin : exp(-2 pi i j / maxrin)
    __shared__ float s_red[kPolarThreads / 32][2 * RPB];
    __shared__ Chunk s_chunk;
    __shared__ float s_cx[RPB], s_cy[RPB], s_avg[RPB], s_isg[RPB];
    __shared__ int4 s_ring[CRA_MAX_RINGS];

    const int tid = threadIdx.x;
    if (tid == 0) {
        Chunk c;
        if (MODE == 1) {                       // reference j at the image centre, one row
            c.part = blockIdx.x; c.row0 = blockIdx.x; c.nrow = 1;
            s_cx[0] = (float)(nx / 2 + 1); s_cy[0] = s_cx[0];
        } else if (map.row_start == nullptr) { // single explicit centre (tests)
            c.part = map.p0; c.row0 = 0; c.nrow = 1; s_cx[0] = fix_cx; s_cy[0] = fix_cy;
        } else {
            const int b = blockIdx.x;
            int lo = 0, hi = map.np;           // last p with chunk_start[p] <= b
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (map.chunk_start[mid] <= b) lo = mid; else hi = mid; }
            const int rs = map.row_start[lo], re = map.row_start[lo + 1];
            const int g = rs / RPB + (b - map.chunk_start[lo]);
            c.part = map.p0 + lo;
            c.row0 = max(rs, g * RPB);
            c.nrow = min(re, (g + 1) * RPB) - c.row0;
            const int4 w = map.win[lo];
            const int wx = w.x + w.y + 1;
            for (int r = 0; r < c.nrow; ++r) {
                const int li = c.row0 + r - rs;
                const int i = li / wx - w.z;   // y index, outer loop of multiref_polar_ali_2d
                const int j = li % wx - w.x;   // x index, inner loop
                s_cx[r] = map.search[lo].cx + j * map.step;
                s_cy[r] = map.search[lo].cy + i * map.step;
            }
        }
        s_chunk = c;
    }
    for (int i = tid; i < maxrin; i += kPolarThreads) s_tw[i] = twid[i];
    for (int i = tid; i < tab->nring; i += kPolarThreads) {
        const int n = tab->len[i] >> 1, lg = 31 - __clz(n);
        s_ring[i] = make_int4(tab->poff[i], lg - (lg >> 1), tab->len[i] >> 2, __float_as_int(tab->wn[i]));
    }
    __syncthreads();
    const Chunk ck = s_chunk;
    const float* img = images + (size_t)ck.part * npix;
    if (!GIMG) {
        if ((npix & 3) == 0) {
            const float4* g4 = reinterpret_cast<const float4*>(img);
            float4* s4 = reinterpret_cast<float4*>(s_img);
            for (int i = tid; i < (npix >> 2); i += kPolarThreads) s4[i] = __ldg(g4 + i);
        } else {
            for (int i = tid; i < npix; i += kPolarThreads) s_img[i] = __ldg(img + i);
        }
        __syncthreads();
    }

    // ---- resample every row of the sub-group -------------------------------------------------
    // The table holds one quarter of every ring (alrl_ms builds the other three by symmetry:
    // (x,y) -> (y,-x) -> (-x,-y) -> (-y,x)); 23 KB for ou=36, so it stays L1-resident.
    const float rmax = (float)tab->rad[tab->nring - 1];
    float av[RPB], sq[RPB];
    bool all_inside = true;
#pragma unroll
    for (int r = 0; r < RPB; ++r) {
        av[r] = 0.f; sq[r] = 0.f;
        if (r < ck.nrow) {
            const float cx = s_cx[r], cy = s_cy[r];
            all_inside = all_inside && (cx - rmax >= 2.0f) && (cx + rmax <= (float)(nx - 1)) &&
                         (cy - rmax >= 2.0f) && (cy + rmax <= (float)(nx - 1));
        }
    }
    const int nq = lcirc >> 2;
    for (int q = tid; q < nq; q += kPolarThreads) {
        const float4 e = __ldg(samp + q);                 // x, y, ring, jt
        const int4 rp = s_ring[__float_as_int(e.z)];      // poff, log2 NB, len/4, Normalize_ring weight
        const int jt = __float_as_int(e.w);
        const float wn = __int_as_float(rp.w);
        const float ox[4] = {e.x, e.y, -e.x, -e.y}, oy[4] = {e.y, -e.x, -e.y, e.x};
        int slot[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int j = jt + m * rp.z, p = j >> 1;
            slot[m] = 2 * (rp.x + p + (p >> rp.y)) + (j & 1);
        }
#pragma unroll
        for (int r = 0; r < RPB; ++r) {
            if (r < ck.nrow) {
                const float cx = s_cx[r], cy = s_cy[r];
                float* circ = s_circ + r * lcp;
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    float v;
                    if (GIMG) v = all_inside ? quadri_inside(ox[m] + cx, oy[m] + cy, nx, img) : quadri_smem(ox[m] + cx, oy[m] + cy, nx, img);
                    else      v = all_inside ? quadri_inside(ox[m] + cx, oy[m] + cy, nx, s_img) : quadri_smem(ox[m] + cx, oy[m] + cy, nx, s_img);
                    circ[slot[m]] = v;
                    if (MODE == 0) { av[r] += v * wn; sq[r] += v * v * wn; }
                }
            }
        }
    }
    if (MODE == 0 && normalize_ring) {
#pragma unroll
        for (int r = 0; r < RPB; ++r) { av[r] = warp_sum(av[r]); sq[r] = warp_sum(sq[r]); }
        if ((tid & 31) == 0) {
#pragma unroll
            for (int r = 0; r < RPB; ++r) { s_red[tid >> 5][2 * r] = av[r]; s_red[tid >> 5][2 * r + 1] = sq[r]; }
        }
    }
    __syncthreads();
    if (tid < RPB) {
        float avg = 0.f, isg = 1.f;
        if (MODE == 0 && normalize_ring) {
            float a = 0.f, s = 0.f;
#pragma unroll
            for (int w = 0; w < kPolarThreads / 32; ++w) { a += s_red[w][2 * tid]; s += s_red[w][2 * tid + 1]; }
            const float nn = tab->nn;
            avg = a / nn;
            isg = 1.0f / sqrtf((s - a * a / nn) / nn);
        }
        s_avg[tid] = avg; s_isg[tid] = isg;
    }

#ifdef CRA_EXP_SKIP_RINGFFT
    return;
#endif
    // ---- ring FFTs: pass A -------------------------------------------------------------------
    for (int it = tid; it < items.nA * ck.nrow; it += kPolarThreads) {
        const int r = it / items.nA, item = __ldg(items.A + (it - r * items.nA));
        const int ring = item >> 16, b = item & 0xffff;
        const int len = tab->len[ring];
        float2* z = reinterpret_cast<float2*>(s_circ + r * lcp) + tab->poff[ring];
        const int lg = 31 - __clz(len >> 1);
        const float2 base = s_tw[b * (maxrin / (len >> 1))];     // exp(-2 pi i b / n)
        switch (lg) {
            case 2: pass_a<2, 2>(z, b, base); break;
            case 3: pass_a<2, 4>(z, b, base); break;
            case 4: pass_a<4, 4>(z, b, base); break;
            case 5: pass_a<4, 8>(z, b, base); break;
            case 6: pass_a<8, 8>(z, b, base); break;
            case 7: pass_a<8, 16>(z, b, base); break;
            case 8: pass_a<16, 16>(z, b, base); break;
            default: pass_a<16, 32>(z, b, base); break;
        }
    }
    __syncthreads();
    // ---- pass B ------------------------------------------------------------------------------
    for (int it = tid; it < items.nB * ck.nrow; it += kPolarThreads) {
        const int r = it / items.nB, item = __ldg(items.B + (it - r * items.nB));
        const int ring = item >> 16, ka = item & 0xffff;
        const int len = tab->len[ring];
        float2* z = reinterpret_cast<float2*>(s_circ + r * lcp) + tab->poff[ring];
        const int lg = 31 - __clz(len >> 1);
        switch (lg) {
            case 2: pass_b<2, 2>(z, ka); break;
            case 3: pass_b<2, 4>(z, ka); break;
            case 4: pass_b<4, 4>(z, ka); break;
            case 5: pass_b<4, 8>(z, ka); break;
            case 6: pass_b<8, 8>(z, ka); break;
            case 7: pass_b<8, 16>(z, ka); break;
            case 8: pass_b<16, 16>(z, ka); break;
            default: pass_b<16, 32>(z, ka); break;
        }
    }
    __syncthreads();
    // ---- pass C: real-FFT split, Normalize_ring / Applyws, store ------------------------------
    // Z_k of the half-length complex FFT sits at z[(k % NA)*(NB+1) + k / NA];
    // F_k = E_k + w_k O_k, F_{n-k} = conj(E_k - w_k O_k), w_k = exp(-2 pi i k / len)
    const bool vec = (ck.nrow == RPB) && (ck.row0 % RPB == 0);
    float2* const grp = reinterpret_cast<float2*>(spec) + (size_t)(ck.row0 >> 2) * tab->nc * 4;
    for (int it = tid; it < items.nC; it += kPolarThreads) {
        const int item = __ldg(items.C + it);
        const int ring = item >> 16, k = item & 0xffff;
        const int len = tab->len[ring], n = len >> 1;
        const int lg = 31 - __clz(n);
        const int la = lg >> 1, NA = 1 << la, NB = n >> la;
        const int m = (k == 0) ? 0 : n - k;
        const int pk = (k & (NA - 1)) * (NB + 1) + (k >> la), pm = (m & (NA - 1)) * (NB + 1) + (m >> la);
        const float2 wk = s_tw[k * (maxrin / len)];
        float wgt = 1.0f, wnyq = 1.0f;
        if (MODE == 1) { wgt = tab->wr[ring]; wnyq = (len != maxrin) ? 0.5f * wgt : wgt; }
        float2 fk[RPB], fm[RPB];
#pragma unroll
        for (int r = 0; r < RPB; ++r) {
            fk[r] = make_float2(0.f, 0.f); fm[r] = fk[r];
            if (r < ck.nrow) {
                const float2* z = reinterpret_cast<const float2*>(s_circ + r * lcp) + tab->poff[ring];
                const float2 a = z[pk], b = z[pm];
                const float sc = s_isg[r];
                if (k == 0) {
                    float dc = a.x + a.y, ny = a.x - a.y;
                    if (MODE == 0) { dc = (dc - s_avg[r] * (float)len) * sc; ny *= sc; }
                    fk[r] = make_float2(dc * wgt, 0.f);
                    fm[r] = make_float2(ny * wnyq, 0.f);           // F_n: the ring's Nyquist term
                } else {
                    const float2 E = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
                    const float2 O = make_float2(0.5f * (a.y + b.y), -0.5f * (a.x - b.x));
                    const float2 P = cmul(O, wk);
                    const float s2 = (MODE == 0) ? sc : wgt;
                    fk[r] = make_float2((E.x + P.x) * s2, (E.y + P.y) * s2);
                    fm[r] = make_float2((E.x - P.x) * s2, -(E.y - P.y) * s2);
                }
            }
        }
        if (FMT == CRA_FMT_FRAG) {
            // back into the ring buffer in place: slot pos(k) <- F_k, pos(n-k) <- F_{n-k}; pos(0) <- (F_0, F_n)
#pragma unroll
            for (int r = 0; r < RPB; ++r)
                if (r < ck.nrow) {
                    float2* z = reinterpret_cast<float2*>(s_circ + r * lcp) + tab->poff[ring];
                    if (k == 0) z[pk] = make_float2(fk[r].x, fm[r].x);
                    else { z[pk] = fk[r]; if (pm != pk) z[pm] = fm[r]; }
                }
            continue;
        }
        const int coff = tab->coff[ring], km = (k == 0) ? n : m;     // n = len/2: the Nyquist slot
        if (vec && RPB == 4) {
            float4* o0 = reinterpret_cast<float4*>(grp) + 2 * coff, *o1 = o0 + (n + 1);
            o0[k] = make_float4(fk[0].x, fk[0].y, fk[1 % RPB].x, fk[1 % RPB].y);
            o1[k] = make_float4(fk[2 % RPB].x, fk[2 % RPB].y, fk[3 % RPB].x, fk[3 % RPB].y);
            if (km != k) {
                o0[km] = make_float4(fm[0].x, fm[0].y, fm[1 % RPB].x, fm[1 % RPB].y);
                o1[km] = make_float4(fm[2 % RPB].x, fm[2 % RPB].y, fm[3 % RPB].x, fm[3 % RPB].y);
            }
        } else if (vec && RPB == 2) {
            float4* o = reinterpret_cast<float4*>(grp) + 2 * coff + ((ck.row0 & 3) >> 1) * (n + 1);
            o[k] = make_float4(fk[0].x, fk[0].y, fk[1 % RPB].x, fk[1 % RPB].y);
            if (km != k) o[km] = make_float4(fm[0].x, fm[0].y, fm[1 % RPB].x, fm[1 % RPB].y);
        } else {
#pragma unroll
            for (int r = 0; r < RPB; ++r)
                if (r < ck.nrow) {
                    grp[cra_spec_idx(coff, n, ck.row0 + r, k)] = fk[r];
                    if (km != k) grp[cra_spec_idx(coff, n, ck.row0 + r, km)] = fm[r];
                }
        }
    }
    if (FMT == CRA_FMT_FRAG) {
        // ---- pass D: gather the 4 ring slots of every (chunk, quad) and store the split-bf16 units ----
        __syncthreads();
        const int nch = frag.nch, nring = tab->nring;
        const int per_row = nch * 4;
        unsigned char* const base = reinterpret_cast<unsigned char*>(spec) + (size_t)ck.row0 * cra_frag_row_bytes(nch);
        for (int it = tid; it < per_row * ck.nrow; it += kPolarThreads) {
            const int r = it / per_row, u = it - r * per_row;
            const int gc = u >> 2, t = u & 3;
            const int kc = __ldg(frag.chunk_k + gc);
            const int k = kc >> 4, s0 = 16 * (kc & 15) + 4 * t;
            const float2* zrow = reinterpret_cast<const float2*>(s_circ + r * lcp);
            float re[4], im[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ring = nring - 1 - (s0 + j);
                float2 v = make_float2(0.f, 0.f);
                if (ring >= 0) {
                    const int4 rp = s_ring[ring];
                    const int n = rp.z * 2;                      // len/2
                    if (k <= n) {
                        const float2* z = zrow + rp.x;
                        const int la = (31 - __clz(n)) - rp.y;   // log2 NA
                        if (k == 0) v = make_float2(z[0].x, 0.f);
                        else if (k == n) v = make_float2(z[0].y, 0.f);
                        else v = z[(k & ((1 << la) - 1)) * ((1 << rp.y) + 1) + (k >> la)];
                    }
                }
                re[j] = v.x; im[j] = v.y;
            }
            uint4* o = reinterpret_cast<uint4*>(base + (size_t)r * cra_frag_row_bytes(nch) + (size_t)gc * 128 + t * 32);
            uint4 a, b;
            if (MODE == 1 || frag.unit_rows) {           // reference (B operand) layout: [re unit | im unit]
                a = split_bf16x4(re[0], re[1], re[2], re[3]);
                b = split_bf16x4(im[0], im[1], im[2], im[3]);
            } else split_row_unit(re, im, a, b);         // particle row (A operand) layout: hi words then lo words
            // one 256-bit store per lane (the lanes of a warp write different lines: store count loads the LSU)
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         :: "l"(o), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
        }
        if (MODE == 0 && norm != nullptr && tid < ck.nrow) norm[ck.row0 + tid] = make_float2(0.f, 1.f);   // normalised above
        if (MODE == 1 && tref != nullptr && tid < 32) {
            // tref = sum over rings of len * (weighted DC), fixed summation order (cra_common.cuh)
            const float2* zrow = reinterpret_cast<const float2*>(s_circ);
            float a = 0.f;
            for (int i = tid; i < nring; i += 32) a += (float)tab->len[i] * zrow[tab->poff[i]].x;
            a = warp_sum(a);
            if (tid == 0) tref[ck.row0] = a;
        }
    }
}

// normalize.mask: mode 0 -> x - mean_mask ; mode 1 -> (x - mean_mask)/sigma_mask(n-1) ; mode 2 -> image untouched.
// dc_out (optional): the in-mask mean the image still carries afterwards (mode 2: the mean; else 0) -- the grouped
// row kernel removes it before the ring FFTs when Normalize_ring is on (cra_polar_grp.cu).
__global__ void __launch_bounds__(256)
mask_normalize_kernel(float* __restrict__ imgs, int npix, const float* __restrict__ mask, int mode, float* __restrict__ dc_out)
{
    float* img = imgs + (size_t)blockIdx.x * npix;
    double sum = 0.0, sq2 = 0.0; int cnt = 0;
    for (int i = threadIdx.x; i < npix; i += blockDim.x)
        if (mask[i] > 0.5f) { float v = img[i]; sum += v; sq2 += (double)v * v; ++cnt; }
    __shared__ double s_a[8], s_b[8]; __shared__ int s_c[8];
    __shared__ float s_mean, s_sig;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq2 += __shfl_xor_sync(0xffffffffu, sq2, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = sum; s_b[threadIdx.x >> 5] = sq2; s_c[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0; int c = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_a[w]; b += s_b[w]; c += s_c[w]; }
        s_mean = (c == 0) ? 0.f : (float)a / (float)c;
        s_sig = (mode != 1) ? 1.0f : sqrtf((float)((b - a * a / c) / (c - 1)));
    }
    __syncthreads();
    const float mean = s_mean, sig = s_sig;
    if (dc_out && threadIdx.x == 0) dc_out[blockIdx.x] = (mode == 2) ? mean : 0.f;
    if (mode == 2) return;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) img[i] = (img[i] - mean) / sig;
}

size_t polar_smem_bytes(int rpb, bool gimg, int nx, const CraRingTab& h)
{
    size_t npix = gimg ? 0 : (((size_t)nx * nx + 3) & ~(size_t)3);
    size_t lc = (size_t)h.lcpad;
    return (npix + rpb * lc) * sizeof(float) + (size_t)h.maxrin * sizeof(float2);
}

template <int MODE, int RPB, int FMT, bool GIMG>
int launch_polar_g(const float* images, int nx, const CraRingTab* tab, const CraRingTab& htab,
                   const float4* samp, const float* sampw, const float2* twid, const CraPolarItems& items,
                   CraRowMap map, float cx, float cy, int normalize_ring, float* spec, const CraFragTab& frag,
                   float2* norm, float* tref, int nblocks, cudaStream_t st)
{
    if (nblocks <= 0) return 0;
    size_t smem = polar_smem_bytes(RPB, GIMG, nx, htab);
    if (cra_ensure_dyn_smem(reinterpret_cast<const void*>(&polar_fft_kernel<MODE, RPB, FMT, GIMG>), smem)) return 1;
    polar_fft_kernel<MODE, RPB, FMT, GIMG><<<nblocks, kPolarThreads, smem, st>>>(images, nx, tab, samp, sampw, twid, items, map,
                                                                                cx, cy, normalize_ring, spec, frag, norm, tref);
    CRA_CUDA(cudaGetLastError());
    return 0;
}

template <int MODE, int RPB, int FMT>
int launch_polar_f(const float* images, int nx, const CraRingTab* tab, const CraRingTab& htab,
                   const float4* samp, const float* sampw, const float2* twid, const CraPolarItems& items,
                   CraRowMap map, float cx, float cy, int normalize_ring, float* spec, const CraFragTab& frag,
                   float2* norm, float* tref, int nblocks, bool gimg, cudaStream_t st)
{
    if (gimg)
        return launch_polar_g<MODE, RPB, FMT, true>(images, nx, tab, htab, samp, sampw, twid, items, map, cx, cy, normalize_ring,
                                                    spec, frag, norm, tref, nblocks, st);
    return launch_polar_g<MODE, RPB, FMT, false>(images, nx, tab, htab, samp, sampw, twid, items, map, cx, cy, normalize_ring,
                                                 spec, frag, norm, tref, nblocks, st);
}

template <int MODE, int RPB>
int launch_polar(const float* images, int nx, const CraRingTab* tab, const CraRingTab& htab,
                 const float4* samp, const float* sampw, const float2* twid, const CraPolarItems& items,
                 CraRowMap map, float cx, float cy, int normalize_ring, float* spec, int fmt, const CraFragTab& frag,
                 float2* norm, float* tref, int nblocks, bool gimg, cudaStream_t st)
{
    if (fmt == CRA_FMT_FRAG)
        return launch_polar_f<MODE, RPB, CRA_FMT_FRAG>(images, nx, tab, htab, samp, sampw, twid, items, map, cx, cy,
                                                       normalize_ring, spec, frag, norm, tref, nblocks, gimg, st);
    return launch_polar_f<MODE, RPB, CRA_FMT_F32>(images, nx, tab, htab, samp, sampw, twid, items, map, cx, cy,
                                                  normalize_ring, spec, frag, nullptr, nullptr, nblocks, gimg, st);
}

}  // namespace

// Which variant of the general kernel a geometry gets: rows per CTA and whether the image is staged in shared memory.
// Preference: RPB rows + image tile, one row + image tile, RPB rows from global memory, one row from global memory.
// Returns 1 when not even one polar row fits (ou beyond ~108 with maxrin 1024).
int cra_polar_general_layout(int nx, const CraRingTab& htab, size_t smem_limit, int want_rpb, int* rpb, int* gimg)
{
    const int cand_rpb[2] = {want_rpb, 1};
    for (int g = 0; g < 2; ++g)
        for (int i = 0; i < 2; ++i)
            if (polar_smem_bytes(cand_rpb[i], g != 0, nx, htab) + 4096 <= smem_limit) { *rpb = cand_rpb[i]; *gimg = g; return 0; }
    return 1;
}
int cra_polar_default_rpb() { return CRA_POLAR_RPB; }

int cra_launch_mask_normalize(float* imgs, int n, int nx, const float* mask, int mode, float* dc_out, cudaStream_t st)
{
    if (n <= 0) return 0;
    mask_normalize_kernel<<<n, 256, 0, st>>>(imgs, nx * nx, mask, mode, dc_out);
    CRA_CUDA(cudaGetLastError());
    return 0;
}

int cra_launch_polar_rows(const float* images, int nx, const CraRingTab* tab, const CraRingTab& htab,
                          const float4* samp, const float* sampw, const float2* twid, const CraPolarItems& items,
                          CraRowMap map, int normalize_ring, float* spec, int fmt, const CraFragTab& frag,
                          float2* norm, int rpb, int gimg, cudaStream_t st)
{
    if (rpb == 1)
        return launch_polar<0, 1>(images, nx, tab, htab, samp, sampw, twid, items, map, 0.f, 0.f,
                                  normalize_ring, spec, fmt, frag, norm, nullptr, map.nchunks, gimg != 0, st);
    if (rpb != CRA_POLAR_RPB) { cra_set_error("general row kernel: unsupported rows per block"); return 1; }
    return launch_polar<0, CRA_POLAR_RPB>(images, nx, tab, htab, samp, sampw, twid, items, map, 0.f, 0.f,
                                          normalize_ring, spec, fmt, frag, norm, nullptr, map.nchunks, gimg != 0, st);
}

int cra_launch_polar_refs(const float* refs, int R, int nx, const CraRingTab* tab, const CraRingTab& htab,
                          const float4* samp, const float2* twid, const CraPolarItems& items, float* refspec,
                          int fmt, const CraFragTab& frag, float* tref, int gimg, cudaStream_t st)
{
    CraRowMap map{};
    return launch_polar<1, 1>(refs, nx, tab, htab, samp, nullptr, twid, items, map, 0.f, 0.f, 0, refspec, fmt, frag,
                              nullptr, tref, R, gimg != 0, st);
}

int cra_launch_polar_single(const float* image, int nx, const CraRingTab* tab, const CraRingTab& htab,
                            const float4* samp, const float* sampw, const float2* twid, const CraPolarItems& items,
                            float cx, float cy, int normalize_ring, float* spec, int fmt, const CraFragTab& frag,
                            float2* norm, int gimg, cudaStream_t st)
{
    CraRowMap map{};
    map.row_start = nullptr; map.p0 = 0;
    return launch_polar<0, 1>(image, nx, tab, htab, samp, sampw, twid, items, map, cx, cy, normalize_ring, spec, fmt, frag,
                              norm, nullptr, 1, gimg != 0, st);
}
