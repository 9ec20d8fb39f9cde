// cra_polar_grp.cu -- the production row kernel of stage 1+2: Util.Polar2Dm (alrl_ms + quadri),
// the sums of Normalize_ring and the per-ring real FFT (Util.Frngs) for EVERY shift of a particle
// at once.  Reference call site: test_mref.py:200-201 -> the (iy, ix) double loop of EMAN2
// Util::multiref_polar_ali_2d; replaces the per-shift relaunch of cu_resample_to_polar + cuFFT R2C
// (cuda/gpu_aln_noref.cu:818-879, :1775-1816).
//
// When the search step is a whole number of pixels every shift of a particle samples the image at
// the same sub-pixel phase: the six quadri weights of a ring sample depend on the fractional part
// of (centre + sample offset) only, so they are computed ONCE per sample and applied to up to
// RMAX shift rows, whose taps differ by an integer pixel offset.  One CTA owns (particle, block of
// <= rmax shift rows): the image tile is staged once in shared memory and the CTA walks the rings
// in PHASES of 4-ring units (the unit of the fragment layout, cra_common.cuh) so that the ring
// buffers of all its rows fit beside the image:
//   interpolate (weights shared, Normalize_ring sums in registers) -> ring FFT pass A -> pass B
//   -> real-FFT split (index math shared by the rows) -> split-bf16 + 32-byte unit stores.
// Normalize_ring is deferred (cra_common.cuh): the spectrum of the raw polar image is stored and
// (avg, 1/sigma) goes to norm[row]; the CCF kernel applies it when it emits a candidate.
// The shared-memory tile is the image itself, dense, filled by ONE bulk asynchronous copy (TMA, cra_tma.cuh)
// that runs under the CTA's table set-up.  search_range keeps every sample in [2, nx] (1-based), so the six taps
// of a sample stay inside the image except for a sample exactly on the last column / row -- which sits on a pixel
// boundary and is therefore redone by repair_sample, with quadri's circular closure written out.
// Requirements checked by the host (cra_api.cu): a step the phase classes cover and that sample range; otherwise
// the general kernel of cra_polar.cu runs.
#include "cra_common.cuh"
#include "cra_fft.cuh"
#include "cra_tma.cuh"
#include <cuda_bf16.h>

namespace {

using crafft::cmul;
using crafft::fft_reg;

constexpr int kThreads = 256;
constexpr int RMAX = CRA_GRP_RMAX;
constexpr int kFragCap = 24;    // per warp and phase; beyond it the lane repairs its sample itself (exact either way)
// The CTA works as NH (1 or 2) independent thread groups: they share the image tile and the tables, each takes an equal
// share of the block's rows through every phase and synchronises on its own named barrier.  Half as many warps wait
// for the slowest one at each of the 42 barriers of a CTA, the groups drift apart so that one's interpolation (LSU)
// runs under the other's FFT passes (FMA), and the per-row register arrays shrink from RMAX to RMAX / NH entries.
// NH is a template parameter chosen by the host (cra_api.cu: build_group_plan): two groups where two CTAs of >= 12 rows
// are resident (rmax = 17 at nx = 90: 135.9 -> 125.7 ms per step); at nx = 128 / ou = 60, where one CTA of 17 rows fills the
// SM and every phase holds twice the samples per thread, two groups measured +12 % and the launcher takes one.
template <int NH>
__device__ __forceinline__ void group_sync(int g)
{
    constexpr int GT = kThreads / NH;
    if (NH == 1) { __syncthreads(); return; }
    // immediate barrier numbers: a register operand makes ptxas reserve all 16 named barriers of the CTA
    if (g == 0)      asm volatile("bar.sync 1, %0;" :: "n"(GT) : "memory");
    else if (g == 1) asm volatile("bar.sync 2, %0;" :: "n"(GT) : "memory");
    else if (g == 2) asm volatile("bar.sync 3, %0;" :: "n"(GT) : "memory");
    else             asm volatile("bar.sync 4, %0;" :: "n"(GT) : "memory");
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

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

// split-bf16 of four values: (hi01, hi23, lo01, lo23), value = hi + lo: one 16-byte unit of the fragment layout
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
// one 32-byte quad of a particle row in the active row layout (cra_common.cuh)
__device__ __forceinline__ void store_row_quad(uint4* o4, const float (&re)[4], const float (&im)[4], int unit_rows)
{
    uint4 a, b;
    if (unit_rows) {
        a = split_bf16x4(re[0], re[1], re[2], re[3]);
        b = split_bf16x4(im[0], im[1], im[2], im[3]);
    } else split_row_unit(re, im, a, b);
    // one 256-bit store per lane: the 32 lanes of a warp write 32 different lines (the units of a chunk belong to
    // different phases), so the store count, not the byte count, is what loads the LSU
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "l"(o4), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

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

// floor(w / d) for w * magic < 2^32, magic = floor(2^24 / d) + 1 (host side, exact for the ranges used here)
__device__ __forceinline__ int fastdiv(int w, int magic) { return (int)(((unsigned)w * (unsigned)magic) >> 24); }

// One fragile sample (code = q * 4 + m) of block row r, positioned exactly as the reference does it
// (x = offset + the row's own centre) and interpolated with quadri's own expression; the Normalize_ring
// sums of the row are corrected through s_fix.
__device__ __forceinline__ void repair_sample(int code, int r, const float4* __restrict__ samp, const int4* s_ring,
                                              const float2* s_rowc, const float* s_img, int nx, int pitch, float* s_buf,
                                              int stride, float* s_fix)
{
    const int q = code >> 2, m = code & 3;
    const float4 e = __ldg(samp + q);
    const int4 rp = s_ring[__float_as_int(e.z)];
    const float wn = __int_as_float(rp.w);
    const float fx = (m == 0) ? e.x : (m == 1) ? e.y : (m == 2) ? -e.x : -e.y;
    const float fy = (m == 0) ? e.y : (m == 1) ? -e.x : (m == 2) ? -e.y : e.x;
    const int j = __float_as_int(e.w) + m * rp.z, pp = j >> 1;
    const int sl = 2 * (rp.x + pp + (pp >> rp.y)) + (j & 1);
    const float2 c = s_rowc[r];
    const float X = fx + c.x, Y = fy + c.y;
    const int ix = (int)X, iy = (int)Y;                       // 1-based cell, 1 <= ix, iy <= nx
    const float dx = X - (float)ix, dy = Y - (float)iy;
    // quadri's circular closure of the neighbours (Util::quadri: ip1 > nx -> ip1 - nx, im1 < 1 -> im1 + nx)
    const int x0 = ix - 1, xp = (ix + 1 > nx) ? 0 : ix, xm = (ix - 1 < 1) ? nx - 1 : ix - 2;
    const int y0 = iy - 1, yp = (iy + 1 > nx) ? 0 : iy, ym = (iy - 1 < 1) ? nx - 1 : iy - 2;
    // s_img: pixel (0, 0) of the image (a windowed tile passes its origin-adjusted base; its samples never wrap)
    const float f0 = s_img[y0 * pitch + x0];
    const float c1 = s_img[y0 * pitch + xp] - f0, c2 = (c1 - f0 + s_img[y0 * pitch + xm]) * 0.5f;
    const float c3 = s_img[yp * pitch + x0] - f0, c4 = (c3 - f0 + s_img[ym * pitch + x0]) * 0.5f;
    const float c5 = s_img[yp * pitch + xp] - f0 - c1 - c3;
    const float v = f0 + dx * (c1 + (dx - 1.0f) * c2 + dy * c5) + dy * (c3 + (dy - 1.0f) * c4);
    float* dst = s_buf + r * stride;
    const float old = dst[sl];
    dst[sl] = v;
    atomicAdd(&s_fix[2 * r], (v - old) * wn);
    atomicAdd(&s_fix[2 * r + 1], (v * v - old * old) * wn);
}

#ifndef CRA_GRP_EXP                 // timing experiments (not valid kernels): 1 no interpolation, 2 no FFT passes, 4 no split, 8 no unit gather / stores
#define CRA_GRP_EXP 0
#endif
#ifndef CRA_GRP_MUNROLL
#define CRA_GRP_MUNROLL 4           // quadrants of a sample quartet unrolled in the interpolation loop (code size against ILP)
#endif
#ifndef CRA_GRP_MINB
#define CRA_GRP_MINB 2              // resident CTAs per SM the register allocation aims for
#endif
// WIN: the tile is a window of the image (plan.tile) instead of the whole image; a template parameter so that the
// whole-image instantiation of the named configurations keeps its constant pitch and tile base
template <int NH, bool WIN>
__global__ void __launch_bounds__(kThreads, CRA_GRP_MINB)
polar_group_kernel(const float* __restrict__ images, int nx, const CraRingTab* __restrict__ tab,
                   const float4* __restrict__ samp, const float2* __restrict__ twid, CraPolarItems items,
                   CraGroupPlan plan, CraRowMap map, int normalize_ring, unsigned char* __restrict__ spec,
                   CraFragTab frag, float2* __restrict__ norm)
{
    static_assert(kThreads % (32 * NH) == 0 && NH >= 1 && NH <= 4, "groups are whole warps");
    constexpr int GT = kThreads / NH;              // threads per group
    constexpr int HR = (RMAX + NH - 1) / NH;       // most rows per group
    extern __shared__ __align__(16) float smem[];
    const int npix = nx * nx;
    const int maxrin = tab->maxrin, nring = tab->nring;
    const int stride = plan.stride;                            // floats per row of the phase buffer
    // dense tile: the whole image (1-based pixel (i, j) at (j - 1) * nx + (i - 1)) or, with plan.tile, a square window of
    // that side around the particle's search window whose origin (s_org) the CTA's first thread places
    const int pitch = WIN ? plan.tile : nx;
    const int tpix = WIN ? plan.tile * plan.tile : npix;
    float* s_img = smem;                                       // pitch * pitch (padded to 4)
    float* s_buf = smem + ((tpix + 3) & ~3);                   // rmax * stride
    float2* s_tw = reinterpret_cast<float2*>(s_buf + plan.rmax * stride);   // maxrin : exp(-2 pi i j / maxrin)
    int4* s_ring = reinterpret_cast<int4*>(s_tw + maxrin);    // nring : phase-local float2 offset, log2 NB, len/4, wn
    int* s_koff = reinterpret_cast<int*>(s_ring + nring);     // maxrin/2 + 2 : first chunk of frequency k
    int* s_unk = s_koff + (maxrin / 2 + 2);                   // units : longest half length of unit u
    __shared__ float s_red[kThreads / 32][2 * HR];
    __shared__ int s_rowoff[RMAX];
    __shared__ int s_grow[RMAX];                               // row of the batch (spectrum / norm index) of block row r
    __shared__ unsigned s_samemask;                            // bit r: row r sits one pixel right of row r-1
    __shared__ float2 s_rowc[RMAX];
    __shared__ float s_fix[2 * RMAX];
    __shared__ int s_frag[kThreads / 32][kFragCap];            // per warp: queued fragile samples of the phase, q * 4 + m
    __shared__ int s_nfrag[kThreads / 32];                          // Normalize_ring sum corrections of the fragile samples                            // the reference's own centre of every row
    __shared__ int s_blk[4];                                   // particle (batch local), first local row, rows
    __shared__ float s_base[2];
    __shared__ int s_cls[4];                                   // phase class of the block: cx, cy, rows per window line, sub
    __shared__ int s_org2[2];                                  // windowed tile: image pixel (x, y) of its first element
    __shared__ __align__(8) unsigned long long s_bar;          // completion of the image tile's bulk copy

    const int tid = threadIdx.x;
    if (tid < 32) {
        // Which particle this CTA belongs to: the last p with chunk_start[p] <= blockIdx.x, by a 32-ary search of
        // warp 0 (three rounds of one load per lane for ~1600 particles, against eleven dependent loads of a binary
        // search: the whole CTA waits for this).
        const int b = blockIdx.x;
        int lo = 0, hi = map.np;                               // chunk_start[lo] <= b < chunk_start[hi]
        while (hi - lo > 1) {
            const int stepw = (hi - lo + 31) >> 5;
            const int pos = lo + (tid + 1) * stepw;
            const bool le = (pos < hi) && (__ldg(map.chunk_start + pos) <= b);
            const int k = __popc(__ballot_sync(0xffffffffu, le));  // the probes are monotone: a prefix answers "<="
            hi = min(lo + (k + 1) * stepw, hi);
            lo = lo + k * stepw;
        }
        if (tid == 0) {
            // the image tile: one bulk asynchronous copy, in flight while the tables below are set up
            const float* img = images + (size_t)(map.p0 + lo) * npix;
            const int4 w = map.win[lo];
            bool bulk;
            if (WIN) {
                // the window every shift row of this particle samples (taps included), origin on a 16-byte boundary and
                // inside the frame (the host admits a windowed batch only if all its samples stay clear of the border)
                const int T = plan.tile;
                const bool al4 = (nx & 3) == 0;                    // lines of the image start on 16-byte boundaries
                const int ro = (int)tab->rad[tab->nring - 1];
                int x0 = (int)floorf(map.search[lo].cx - (float)w.x * map.step) - ro - 2;
                int y0 = (int)floorf(map.search[lo].cy - (float)w.z * map.step) - ro - 2;
                if (al4) x0 &= ~3;
                x0 = max(0, min(x0, nx - T)); y0 = max(0, min(y0, nx - T));
                s_org2[0] = x0; s_org2[1] = y0;
                bulk = al4 && cratma::bulk_ok(img, 16);
                if (bulk) {                                        // one bulk copy per tile line, all completing on one barrier
                    cratma::mbar_init(&s_bar, T);
                    for (int r = 0; r < T; ++r)
                        cratma::bulk_load(s_img + r * T, img + (size_t)(y0 + r) * nx + x0, (unsigned)(T * sizeof(float)), &s_bar);
                }
            } else {
                bulk = cratma::bulk_ok(img, (size_t)npix * sizeof(float));
                if (bulk) {
                    cratma::mbar_init(&s_bar, 1);
                    cratma::bulk_load(s_img, img, (unsigned)(npix * sizeof(float)), &s_bar);
                }
            }
            const int bi = b - map.chunk_start[lo];
            const int wx = w.x + w.y + 1, wy = w.z + w.w + 1;
            // A step of 1/sub pixel (sub = 2, 4) splits the window into sub x sub phase classes: the positions
            // (cx + sub i, cy + sub j) of a class lie one whole pixel apart and share their tap weights.  Blocks
            // never cross a class; with a whole-pixel step there is one class and the rows are consecutive.
            const int sub = cra_group_sub(map.step);
            int cx = 0, cy = 0, ncx = wx, rows_c = wx * wy, nblk_c = 1, left = bi;
            for (int c = 0; c < sub * sub; ++c) {
                cy = c / sub; cx = c - cy * sub;
                ncx = (wx - cx + sub - 1) / sub;
                const int ncy = (wy - cy + sub - 1) / sub;
                rows_c = ncx * ncy;
                nblk_c = (rows_c + plan.rmax - 1) / plan.rmax;
                if (left < nblk_c) break;
                left -= nblk_c;
            }
            const int r_lo = (int)(((long)left * rows_c) / nblk_c), r_hi = (int)(((long)(left + 1) * rows_c) / nblk_c);
            s_blk[0] = lo; s_blk[1] = r_lo; s_blk[2] = r_hi - r_lo; s_blk[3] = (bulk ? 2 : 0) | ((sub == 1) ? 1 : 0);
            s_cls[0] = cx; s_cls[1] = cy; s_cls[2] = ncx; s_cls[3] = sub;
            s_base[0] = map.search[lo].cx + (float)(cx - w.x) * map.step;
            s_base[1] = map.search[lo].cy + (float)(cy - w.z) * map.step;
        }
    }
    for (int i = tid; i < maxrin; i += kThreads) s_tw[i] = twid[i];
    if (tid < 2 * RMAX) s_fix[tid] = 0.f;
    if (tid < kThreads / 32) s_nfrag[tid] = 0;
    for (int i = tid; i < maxrin / 2 + 2; i += kThreads) s_koff[i] = __ldg(frag.koff + i);
    for (int i = tid; i < (nring + 3) / 4; i += kThreads) s_unk[i] = __ldg(plan.unit_nk + i);
    for (int i = tid; i < nring; i += kThreads) {
        const int n = tab->len[i] >> 1, lg = 31 - __clz(n);
        s_ring[i] = make_int4(__ldg(plan.ppoff + i), lg - (lg >> 1), tab->len[i] >> 2, __float_as_int(tab->wn[i]));
    }
    __syncthreads();
    const int nr_all = s_blk[2];
    const int grp = tid / GT, gt = tid - grp * GT;             // thread group and index inside it
    const int rg0 = (grp * nr_all) / NH, nr = ((grp + 1) * nr_all) / NH - rg0;   // the group's rows [rg0, rg0 + nr)
    const bool contig = (s_blk[3] & 1) != 0;                   // block rows are consecutive rows of the batch
    if (tid < 32) {                                            // the block's rows, one lane each (published by the barrier below)
        const int lo = s_blk[0], r = tid, nr = nr_all;
        const int cx = s_cls[0], cy = s_cls[1], ncx = s_cls[2], sub = s_cls[3];
        const int pxs = (sub > 1) ? 1 : (int)map.step;         // pixels between neighbouring rows of a class
        bool cont = false;
        if (r < nr) {
            const int4 w = map.win[lo];
            const int wx = w.x + w.y + 1;
            const int t = s_blk[1] + r, ty = t / ncx, tx = t - ty * ncx;
            const int lix = cx + tx * sub, liy = cy + ty * sub;
            cont = r > 0 && pxs == 1 && tx != 0;               // one pixel right of the previous row
            s_grow[r] = map.row_start[lo] + liy * wx + lix;
            s_rowoff[r] = ty * pxs * pitch + tx * pxs;
            s_rowc[r] = make_float2(map.search[lo].cx + (float)(lix - w.x) * map.step,
                                    map.search[lo].cy + (float)(liy - w.z) * map.step);
        }
        const unsigned same = __ballot_sync(0xffffffffu, cont);
        if (tid == 0) s_samemask = same;
    }
    {
        const float* img = images + (size_t)(map.p0 + s_blk[0]) * npix;
        // A particle uploaded without normalize.mask still carries its in-mask mean.  Normalize_ring cancels any
        // constant exactly in exact arithmetic, but here it is deferred past the split-bf16 contraction, where a large
        // ring DC term would cost digits: remove the constant up front.
        const float dcv = (normalize_ring && map.dc) ? __ldg(map.dc + map.p0 + s_blk[0]) : 0.f;
        if (s_blk[3] & 2) {
            cratma::mbar_wait(&s_bar, 0);                      // the bulk copy has landed
            if (dcv != 0.f)                                    // every waiting thread sees the whole tile
                for (int i = tid; i < tpix; i += kThreads) s_img[i] -= dcv;
        } else if (WIN) {
            const int T = plan.tile, x0 = s_org2[0], y0 = s_org2[1];
            for (int i = tid; i < tpix; i += kThreads) { const int r = i / T, cx = i - r * T; s_img[i] = __ldg(img + (size_t)(y0 + r) * nx + x0 + cx) - dcv; }
        } else {
            for (int i = tid; i < npix; i += kThreads) s_img[i] = __ldg(img + i) - dcv;
        }
    }
    // image pixel (0, 0) as seen through the tile (outside the tile's storage when the tile is a window)
    const float* const s_org = WIN ? s_img - (s_org2[1] * pitch + s_org2[0]) : s_img;
    const float bx = s_base[0], by = s_base[1];
    // whole-pixel centres in every row of the block: the base centre is a whole number and the rows lie whole pixels apart
    const bool cint_x = bx == rintf(bx), cint_y = by == rintf(by);
    float av[HR], sq[HR];
#pragma unroll
    for (int r = 0; r < HR; ++r) { av[r] = 0.f; sq[r] = 0.f; }
    __syncthreads();
    const unsigned samemask = (s_samemask >> rg0) & ~1u;       // bit r: the group's row r continues row r - 1 (never its first)
    // the group's view of the per-row tables and of the row buffers
    const int* const g_rowoff = s_rowoff + rg0;
    const int* const g_grow = s_grow + rg0;
    const float2* const g_rowc = s_rowc + rg0;
    float* const g_fix = s_fix + 2 * rg0;
    float* const g_buf = s_buf + rg0 * stride;
    int* const g_nfrag = s_nfrag + grp * (GT / 32);
    int (*const g_frag)[kFragCap] = s_frag + grp * (GT / 32);

    for (int ph = 0; ph < plan.nphase; ++ph) {
        const CraPhase P = plan.phases[ph];
        // ---- interpolate the phase's rings for every row: weights once per sample ----------------
        for (int q = P.q0 + gt; q < P.q1; q += GT) {
            if (CRA_GRP_EXP & 1) break;
            const float4 e = __ldg(samp + q);                 // x, y, ring, jt (first quarter of the ring)
            const int4 rp = s_ring[__float_as_int(e.z)];
            const int jt = __float_as_int(e.w);
            const float wn = __int_as_float(rp.w);
            float oxm = e.x, oym = e.y;                       // quadrant m of the quartet: (x, y) -> (y, -x) -> (-x, -y) -> (-y, x)
            int fragile = 0;       // samples so close to a pixel boundary that float rounding of the per-row
                                   // position (x = offset + centre, as Polar2Dm forms it) could pick another cell
            int ro[HR];
#pragma unroll
            for (int r = 0; r < HR; ++r) ro[r] = (r < nr) ? g_rowoff[r] : 0;
#if CRA_GRP_MUNROLL == 4
#pragma unroll
#elif CRA_GRP_MUNROLL == 2
#pragma unroll 2
#else
#pragma unroll 1
#endif
            for (int m = 0; m < 4; ++m) {
                const int j = jt + m * rp.z, pj = j >> 1;
                const int slot = 2 * (rp.x + pj + (pj >> rp.y)) + (j & 1);
                const float X = oxm + bx, Y = oym + by;
                const int ix = (int)X, iy = (int)Y;
                const float dx = X - (float)ix, dy = Y - (float)iy;
                // A coordinate that is a whole number because BOTH its ring offset and the block's centre are whole
                // numbers (the axis points of every ring under an integer centre: all of mref_ali2d) is exact in every
                // row -- offset + centre is an integer sum -- so it cannot round into another cell.
                const bool xf = (dx < 1e-4f || dx > 0.9999f) && !(cint_x && oxm == rintf(oxm));
                const bool yf = (dy < 1e-4f || dy > 0.9999f) && !(cint_y && oym == rintf(oym));
                { const float tq = oxm; oxm = oym; oym = -tq; }
                if (xf || yf) fragile |= 1 << m;
                // quadri: f0 + dx (c1 + (dx-1) c2 + dy c5) + dy (c3 + (dy-1) c4) as six tap weights
                const float a2 = 0.5f * dx * (dx - 1.0f), b2 = 0.5f * dy * (dy - 1.0f), ab = dx * dy;
                const float w1 = dx + a2 - ab;          // (i+1, j)
                const float w2 = a2;                    // (i-1, j)
                const float w3 = dy + b2 - ab;          // (i, j+1)
                const float w4 = b2;                    // (i, j-1)
                const float w5 = ab;                    // (i+1, j+1)
                const float w0 = 1.0f - dx - dy - 2.0f * a2 - 2.0f * b2 + ab;
                const float* p0 = s_org + (iy * pitch + ix) - (pitch + 1);     // 1-based cell (ix, iy)
                float* dst = g_buf + slot;
                // a row that continues the window line of the previous one (one pixel to the right) reuses
                // three of its taps: (i-1,j) <- (i,j), (i,j) <- (i+1,j), (i,j+1) <- (i+1,j+1)
                float fl = 0.f, fc = 0.f, fr = 0.f, uc = 0.f, ur = 0.f;
#pragma unroll
                for (int r = 0; r < HR; ++r) {
                    if (r < nr) {
                        const float* p = p0 + ro[r];
                        const bool cont = (samemask >> r) & 1;
                        fl = fc; fc = fr; uc = ur;
                        if (!cont) { fl = p[-1]; fc = p[0]; uc = p[pitch]; }
                        fr = p[1]; ur = p[pitch + 1];
                        float v = w0 * fc;
                        v = fmaf(w1, fr, v);
                        v = fmaf(w2, fl, v);
                        v = fmaf(w3, uc, v);
                        v = fmaf(w4, p[-pitch], v);
                        v = fmaf(w5, ur, v);
                        dst[r * stride] = v;
                        const float tv = v * wn;
                        av[r] += tv; sq[r] = fmaf(tv, v, sq[r]);
                    }
                }
            }
            if (fragile) {         // rare: queue those samples; the warp redoes them together below
#pragma unroll
                for (int m = 0; m < 4; ++m)
                    if (fragile & (1 << m)) {
                        const int at = atomicAdd(&g_nfrag[gt >> 5], 1);
                        if (at < kFragCap) g_frag[gt >> 5][at] = q * 4 + m;
                        else                // queue full (an integer or half-integer centre puts whole rings on cell
                            for (int r = 0; r < nr; ++r)    // boundaries): this lane repairs its own sample right away
                                repair_sample(q * 4 + m, r, samp, s_ring, g_rowc, s_org, nx, pitch, g_buf, stride, g_fix);
                    }
            }
        }
        // ---- fragile samples, row by row exactly as the reference positions them (x = offset + centre);
        // each warp repairs the samples it wrote itself, so no CTA barrier is needed in between ----
        __syncwarp();
        {
            const int wid = gt >> 5;
            const int nf = min(g_nfrag[wid], kFragCap);
            for (int x = gt & 31; x < nf * nr; x += 32) {
                const int en = x / nr, r = x - en * nr;
                repair_sample(g_frag[wid][en], r, samp, s_ring, g_rowc, s_org, nx, pitch, g_buf, stride, g_fix);
            }
        }
        group_sync<NH>(grp);
        // ---- ring FFTs, pass A: (row, ring, column b) flattened ---------------------------------
        {
            const int nA = P.a1 - P.a0;
            for (int w = gt; w < nA * nr; w += GT) {
                if (CRA_GRP_EXP & 2) break;
                const int r = fastdiv(w, P.magicA), item = __ldg(items.A + P.a0 + (w - r * nA));
                const int ring = item >> 16, b = item & 0xffff;
                const int4 rp = s_ring[ring];
                const int half = rp.z * 2;                             // complex points of the ring
                float2* z = reinterpret_cast<float2*>(g_buf + r * stride) + rp.x;
                const int lg = 31 - __clz(half);
                const float2 base = s_tw[b * (maxrin / half)];         // exp(-2 pi i b / n)
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
        }
        group_sync<NH>(grp);
        // ---- pass B ------------------------------------------------------------------------------
        {
            const int nB = P.b1 - P.b0;
            for (int w = gt; w < nB * nr; w += GT) {
                if (CRA_GRP_EXP & 2) break;
                const int r = fastdiv(w, P.magicB), item = __ldg(items.B + P.b0 + (w - r * nB));
                const int ring = item >> 16, ka = item & 0xffff;
                const int4 rp = s_ring[ring];
                float2* z = reinterpret_cast<float2*>(g_buf + r * stride) + rp.x;
                const int lg = 31 - __clz(rp.z * 2);
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
        }
        group_sync<NH>(grp);
        // ---- pass C: real-FFT split in place; index math once per (ring, k), rows inside ---------
        // Z_k of the half-length complex FFT sits at z[(k % NA)*(NB+1) + k / NA];
        // F_k = E_k + w_k O_k, F_{n-k} = conj(E_k - w_k O_k), w_k = exp(-2 pi i k / len);
        // slot pos(0) <- (F_0, F_n): both real.  Main lanes: k = 1 .. n/2 of every ring (list Cg);
        // the k = 0 items are a short flat (ring, row) loop.
        {
            const int nC = P.c1 - P.c0;
            const int nset = __ldg(&plan.phases[ph].nsetC[nr]);
            for (int x = gt; x < nC * nset; x += GT) {
                if (CRA_GRP_EXP & 4) break;
                const int set = fastdiv(x, P.magicC), item = __ldg(items.Cg + P.c0 + (x - set * nC));
                const int ring = item >> 16, k = item & 0xffff;
                const int4 rp = s_ring[ring];
                const int n = rp.z * 2, len = rp.z * 4;
                const int lb = rp.y, NA = n >> lb, la = 31 - __clz(NA), NB1 = (1 << lb) + 1;
                const int m = n - k;
                const int pk = rp.x + (k & (NA - 1)) * NB1 + (k >> la), pm = rp.x + (m & (NA - 1)) * NB1 + (m >> la);
                const float2 wk = s_tw[k * (maxrin / len)];
                float2* z = reinterpret_cast<float2*>(g_buf + set * stride);
                for (int r = set; r < nr; r += nset) {
                    const float2 a = z[pk], b = z[pm];
                    const float2 E = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
                    const float2 O = make_float2(0.5f * (a.y + b.y), -0.5f * (a.x - b.x));
                    const float2 Pk = cmul(O, wk);
                    z[pk] = make_float2(E.x + Pk.x, E.y + Pk.y);
                    if (pm != pk) z[pm] = make_float2(E.x - Pk.x, -(E.y - Pk.y));
                    z += nset * (stride >> 1);
                }
            }
            const int s0 = 4 * P.u0, nsl = min(4 * P.u1, nring) - s0;          // ring slots of the phase
            for (int x = gt; x < nsl * nr; x += GT) {
                const int r = x / nsl, ring = nring - 1 - (s0 + (x - r * nsl));
                float2* z = reinterpret_cast<float2*>(g_buf + r * stride) + s_ring[ring].x;
                const float2 a = z[0];
                z[0] = make_float2(a.x + a.y, a.x - a.y);
            }
        }
        group_sync<NH>(grp);
        // ---- pass D: one (k < longest half length, unit) per lane, rows inside: gather the 4 ring
        // slots, split to bf16 hi/lo, store the 32-byte unit.  The unit's top frequency (real, only
        // its longest rings reach it) is a short flat (unit, row) loop.
        {
            const int upr = P.upr;
            const int nset = __ldg(&plan.phases[ph].nsetD[nr]);
            const size_t rb = cra_frag_row_bytes(frag.nch);
            for (int x = gt; x < upr * nset; x += GT) {
                if (CRA_GRP_EXP & 8) break;
                const int set = fastdiv(x, P.magicD), ditem = __ldg(items.D + P.d0 + (x - set * upr));
                const int k = ditem & 0xffff, u = ditem >> 16;
                // value of slot j = (v.x * mx + v.y * my, v.y * mi) of z[idx]: complex (1,0,1), F_0 = .x of
                // pos(0) (1,0,0), F_n = .y of pos(0) (0,1,0), ring too short or absent (0,0,0)
                int idx[4]; float mx[4], my[4], mi[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ring = nring - 1 - (4 * u + j);
                    idx[j] = 0; mx[j] = 0.f; my[j] = 0.f; mi[j] = 0.f;
                    if (ring >= 0) {
                        const int4 rp = s_ring[ring];
                        const int n = rp.z * 2;
                        if (k <= n) {
                            const int lb = rp.y, NA = n >> lb, la = 31 - __clz(NA);
                            idx[j] = rp.x;
                            if (k == 0) mx[j] = 1.f;
                            else if (k == n) my[j] = 1.f;
                            else { idx[j] = rp.x + (k & (NA - 1)) * ((1 << lb) + 1) + (k >> la); mx[j] = 1.f; mi[j] = 1.f; }
                        }
                    }
                }
                unsigned char* const o0 = spec + (size_t)(s_koff[k] + (u >> 2)) * 128 + (u & 3) * 32;
                unsigned char* o = o0 + (size_t)g_grow[set] * rb;
                const float2* z = reinterpret_cast<const float2*>(g_buf + set * stride);
                for (int r = set; r < nr; r += nset) {
                    float re[4], im[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 v = z[idx[j]];
                        re[j] = fmaf(v.y, my[j], v.x * mx[j]);
                        im[j] = v.y * mi[j];
                    }
                    store_row_quad(reinterpret_cast<uint4*>(o), re, im, frag.unit_rows);
                    if (contig) o += (size_t)nset * rb;
                    else if (r + nset < nr) o = o0 + (size_t)g_grow[r + nset] * rb;
                    z += nset * (stride >> 1);
                }
            }
            const int nu = P.u1 - P.u0;
            for (int x = gt; x < nu * nr; x += GT) {
                const int r = x / nu, u = P.u0 + (x - r * nu);
                const int k = s_unk[u];                                        // the unit's longest half length
                const float2* z = reinterpret_cast<const float2*>(g_buf + r * stride);
                float re[4], im[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ring = nring - 1 - (4 * u + j);
                    re[j] = 0.f; im[j] = 0.f;
                    if (ring >= 0) { const int4 rp = s_ring[ring]; if (rp.z * 2 == k) re[j] = z[rp.x].y; }
                }
                store_row_quad(reinterpret_cast<uint4*>(spec + (size_t)g_grow[r] * rb + (size_t)(s_koff[k] + (u >> 2)) * 128 + (u & 3) * 32),
                               re, im, frag.unit_rows);
            }
        }
        if (gt < GT / 32) g_nfrag[gt] = 0;             // read above before two barriers, next written after this phase's last one
        group_sync<NH>(grp);
    }

    // ---- Normalize_ring sums -> norm[row] = (avg, 1/sigma) ---------------------------------------
#pragma unroll
    for (int r = 0; r < HR; ++r) { av[r] = warp_sum(av[r]); sq[r] = warp_sum(sq[r]); }
    if ((tid & 31) == 0) {
#pragma unroll
        for (int r = 0; r < HR; ++r) { s_red[tid >> 5][2 * r] = av[r]; s_red[tid >> 5][2 * r + 1] = sq[r]; }
    }
    __syncthreads();
    if (gt < nr) {                                             // one thread per row of its group
        float avg = 0.f, isg = 1.f;
        if (normalize_ring) {
            float a = g_fix[2 * gt], s = g_fix[2 * gt + 1];
#pragma unroll
            for (int w = 0; w < GT / 32; ++w) { a += s_red[grp * (GT / 32) + w][2 * gt]; s += s_red[grp * (GT / 32) + w][2 * gt + 1]; }
            const float nn = tab->nn;
            avg = a / nn;
            isg = 1.0f / sqrtf((s - a * a / nn) / nn);
        }
        norm[g_grow[gt]] = make_float2(avg, isg);
    }
}

}  // namespace

size_t cra_polar_group_smem(int nx, int maxrin, const CraGroupPlan& plan)
{
    const size_t side = plan.tile ? (size_t)plan.tile : (size_t)nx;
    const size_t npix = (side * side + 3) & ~(size_t)3;                     // the dense image tile (whole image or window)
    // + twiddles, ring table (<= CRA_MAX_RINGS int4), chunk offsets, unit lengths
    return (npix + (size_t)plan.rmax * plan.stride) * sizeof(float) + (size_t)maxrin * sizeof(float2)
         + (size_t)plan.nring * sizeof(int4) + (size_t)(maxrin / 2 + 2 + (plan.nring + 3) / 4) * sizeof(int);
}

int cra_launch_polar_group(const float* images, int nx, const CraRingTab* tab, const CraRingTab& htab,
                           const float4* samp, const float2* twid, const CraPolarItems& items, const CraGroupPlan& plan,
                           CraRowMap map, int normalize_ring, float* spec, const CraFragTab& frag, float2* norm, cudaStream_t st)
{
    if (map.nchunks <= 0) return 0;
    const size_t smem = cra_polar_group_smem(nx, htab.maxrin, plan);
#define CRA_GRP_LAUNCH(NH_, WIN_)                                                                                         \
    { if (cra_ensure_dyn_smem(reinterpret_cast<const void*>(&polar_group_kernel<NH_, WIN_>), smem)) return 1;             \
      polar_group_kernel<NH_, WIN_><<<map.nchunks, kThreads, smem, st>>>(images, nx, tab, samp, twid, items, plan, map,     \
                                                                        normalize_ring, reinterpret_cast<unsigned char*>(spec), frag, norm); }
    if (plan.nh == 2) { if (plan.tile) CRA_GRP_LAUNCH(2, true) else CRA_GRP_LAUNCH(2, false) }
    else              { if (plan.tile) CRA_GRP_LAUNCH(1, true) else CRA_GRP_LAUNCH(1, false) }
#undef CRA_GRP_LAUNCH
    CRA_CUDA(cudaGetLastError());
    return 0;
}
