// cra_polar.cu -- stage 1+2 of the hot path: batched polar resampler
// (Util.Polar2Dm / alrl_ms + quadri), Normalize_ring, per-ring real FFT (Util.Frngs)
// and, for references, Applyws.  Reference call sites: test_mref.py:172-174 (refs),
// :200-201 -> EMAN2 Util::multiref_polar_ali_2d inner loop (particles).
// Replaces cu_resample_to_polar + cuFFT R2C (cuda/gpu_aln_noref.cu:818-879, :1816).
//
// One CTA per row (= one particle at one shift, or one reference).  The image tile
// is staged in shared memory with 128-bit coalesced loads, the 6-tap quadratic
// interpolation gathers from shared memory, ring sums are reduced with warp
// shuffles, and every ring is transformed in shared memory by one warp.
#include "cra_common.cuh"

namespace {

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

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-place packed real FFT of one ring (len floats at c, 8-byte aligned) by one warp.
// tw[j] = exp(-2 pi i j / maxrin), j < maxrin/2.
__device__ void ring_rfft_warp(float* c, int len, int maxrin, const float2* __restrict__ tw, int lane)
{
    float2* z = reinterpret_cast<float2*>(c);
    const int n = len >> 1;
    const int lg = 31 - __clz(n);
    for (int i = lane; i < n; i += 32) {
        int j = (int)(__brev((unsigned)i) >> (32 - lg));
        if (i < j) { float2 t = z[i]; z[i] = z[j]; z[j] = t; }
    }
    __syncwarp();
    for (int len2 = 2; len2 <= n; len2 <<= 1) {
        const int half = len2 >> 1;
        const int tstep = maxrin / len2;
        for (int b = lane; b < (n >> 1); b += 32) {
            int k = b & (half - 1);
            int s = ((b - k) << 1) + k;
            int e = s + half;
            float2 w = tw[k * tstep];
            float2 u = z[s], v = cmul(z[e], w);
            z[s] = make_float2(u.x + v.x, u.y + v.y);
            z[e] = make_float2(u.x - v.x, u.y - v.y);
        }
        __syncwarp();
    }
    // split: F_k = E_k + w_k O_k, F_{n-k} = conj(E_k - w_k O_k)
    const int tstep = maxrin / len;
    for (int k = 1 + lane; k <= (n >> 1); k += 32) {
        int m = n - k;
        float2 a = z[k], b = z[m];
        float2 E = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
        float2 O = make_float2(0.5f * (a.y + b.y), -0.5f * (a.x - b.x));
        float2 P = cmul(O, tw[k * tstep]);
        z[k] = make_float2(E.x + P.x, E.y + P.y);
        z[m] = make_float2(E.x - P.x, -(E.y - P.y));
    }
    if (lane == 0) {
        float2 a = z[0];
        z[0] = make_float2(a.x + a.y, a.x - a.y);
    }
    __syncwarp();
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// mode: 0 = particle rows (optional Normalize_ring), 1 = references (Applyws)
template <int MODE>
__global__ void __launch_bounds__(kPolarThreads)
polar_fft_kernel(const float* __restrict__ images, int nx, const CraRingTab* __restrict__ tab,
                 const float2* __restrict__ samp, const float* __restrict__ sampw,
                 const float2* __restrict__ twid, CraRowMap map, float fix_cx, float fix_cy,
                 int normalize_ring, float* __restrict__ spec)
{
    extern __shared__ __align__(16) float smem[];
    const int npix = nx * nx;
    const int lcirc = tab->lcirc;
    const int maxrin = tab->maxrin;
    float* s_img = smem;                                     // npix (padded to 4)
    float* s_circ = smem + ((npix + 3) & ~3);                // lcirc
    float2* s_tw = reinterpret_cast<float2*>(s_circ + ((lcirc + 3) & ~3));   // maxrin/2
    __shared__ float s_red[2][kPolarThreads / 32];
    __shared__ int s_part;
    __shared__ float s_cx, s_cy;

    const int row = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid == 0) {
        if (MODE == 1) {                       // reference j at the image centre
            s_part = row; s_cx = (float)(nx / 2 + 1); s_cy = s_cx;
        } else if (map.row_start == nullptr) { // single explicit centre (tests)
            s_part = map.p0; s_cx = fix_cx; s_cy = fix_cy;
        } else {
            int lo = 0, hi = map.np;           // last p with row_start[p] <= row
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (map.row_start[mid] <= row) lo = mid; else hi = mid; }
            int li = row - map.row_start[lo];
            int4 w = map.win[lo];
            int wx = w.x + w.y + 1;
            int i = li / wx - w.z;             // y index, outer loop of multiref_polar_ali_2d
            int j = li % wx - w.x;             // x index, inner loop
            float iy = i * map.step, ix = j * map.step;
            s_part = map.p0 + lo;
            s_cx = map.search[lo].cx + ix;
            s_cy = map.search[lo].cy + iy;
        }
    }
    for (int i = tid; i < maxrin / 2; i += kPolarThreads) s_tw[i] = twid[i];
    __syncthreads();
    const float* img = images + (size_t)s_part * npix;
    if ((npix & 3) == 0) {
        const float4* g4 = reinterpret_cast<const float4*>(img);
        float4* s4 = reinterpret_cast<float4*>(s_img);
        for (int i = tid; i < (npix >> 2); i += kPolarThreads) s4[i] = __ldg(g4 + i);
    } else {
        for (int i = tid; i < npix; i += kPolarThreads) s_img[i] = __ldg(img + i);
    }
    __syncthreads();

    const float cx = s_cx, cy = s_cy;
    float av = 0.f, sq = 0.f;
    for (int i = tid; i < lcirc; i += kPolarThreads) {
        float2 p = samp[i];
        float v = quadri_smem(p.x + cx, p.y + cy, nx, s_img);
        s_circ[i] = v;
        if (MODE == 0) { float w = sampw[i]; av += v * w; sq += v * v * w; }
    }
    if (MODE == 0 && normalize_ring) {
        av = warp_sum(av); sq = warp_sum(sq);
        if ((tid & 31) == 0) { s_red[0][tid >> 5] = av; s_red[1][tid >> 5] = sq; }
        __syncthreads();
        float a = 0.f, s = 0.f;
#pragma unroll
        for (int w = 0; w < kPolarThreads / 32; ++w) { a += s_red[0][w]; s += s_red[1][w]; }
        const float nn = tab->nn;
        const float avg = a / nn;
        const float sgm = sqrtf((s - a * a / nn) / nn);
        for (int i = tid; i < lcirc; i += kPolarThreads) s_circ[i] = (s_circ[i] - avg) / sgm;
    }
    __syncthreads();

    // one warp per ring, longest rings first
    const int warp = tid >> 5, lane = tid & 31, nwarp = kPolarThreads / 32;
    for (int r = tab->nring - 1 - warp; r >= 0; r -= nwarp) {
        const int len = tab->len[r];
        float* c = s_circ + tab->off[r];
        ring_rfft_warp(c, len, maxrin, s_tw, lane);
        if (MODE == 1) {
            const float w = tab->wr[r];
            for (int i = lane; i < len; i += 32) {
                float ww = (i == 1 && len != maxrin) ? 0.5f * w : w;
                c[i] *= ww;
            }
        }
    }
    __syncthreads();
    float* out = spec + (size_t)row * lcirc;
    if ((lcirc & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(s_circ);
        float4* o4 = reinterpret_cast<float4*>(out);
        for (int i = tid; i < (lcirc >> 2); i += kPolarThreads) o4[i] = s4[i];
    } else {
        for (int i = tid; i < lcirc; i += kPolarThreads) out[i] = s_circ[i];
    }
}

// normalize.mask: mode 0 -> x - mean_mask ; mode 1 -> (x - mean_mask)/sigma_mask(n-1)
__global__ void __launch_bounds__(256)
mask_normalize_kernel(float* __restrict__ imgs, int npix, const float* __restrict__ mask, int mode)
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
        s_sig = (mode == 0) ? 1.0f : sqrtf((float)((b - a * a / c) / (c - 1)));
    }
    __syncthreads();
    const float mean = s_mean, sig = s_sig;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) img[i] = (img[i] - mean) / sig;
}

size_t polar_smem_bytes(int nx, const CraRingTab& h)
{
    size_t npix = ((size_t)nx * nx + 3) & ~(size_t)3;
    size_t lc = ((size_t)h.lcirc + 3) & ~(size_t)3;
    return (npix + lc) * sizeof(float) + (size_t)(h.maxrin / 2) * sizeof(float2);
}

}  // namespace

int cra_launch_mask_normalize(float* imgs, int n, int nx, const float* mask, int mode, cudaStream_t st)
{
    if (n <= 0) return 0;
    mask_normalize_kernel<<<n, 256, 0, st>>>(imgs, nx * nx, mask, mode);
    CRA_CUDA(cudaGetLastError());
    return 0;
}

template <int MODE>
static int launch_polar(const float* images, int nx, const CraRingTab* tab, const CraRingTab& htab,
                        const float2* samp, const float* sampw, const float2* twid,
                        CraRowMap map, float cx, float cy, int normalize_ring, float* spec, int nblocks, cudaStream_t st)
{
    if (nblocks <= 0) return 0;
    size_t smem = polar_smem_bytes(nx, htab);
    static size_t configured = 0;
    if (smem > configured) {
        CRA_CUDA(cudaFuncSetAttribute(polar_fft_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    polar_fft_kernel<MODE><<<nblocks, kPolarThreads, smem, st>>>(images, nx, tab, samp, sampw, twid, map,
                                                                cx, cy, normalize_ring, spec);
    CRA_CUDA(cudaGetLastError());
    return 0;
}

int cra_launch_polar_rows(const float* images, int nx, const CraRingTab* tab, const CraRingTab& htab,
                          const float2* samp, const float* sampw, const float2* twid, CraRowMap map,
                          int normalize_ring, float* spec, cudaStream_t st)
{
    return launch_polar<0>(images, nx, tab, htab, samp, sampw, twid, map, 0.f, 0.f, normalize_ring, spec, map.nrows, st);
}

int cra_launch_polar_refs(const float* refs, int R, int nx, const CraRingTab* tab, const CraRingTab& htab,
                          const float2* samp, const float2* twid, float* refspec, cudaStream_t st)
{
    CraRowMap map{};
    return launch_polar<1>(refs, nx, tab, htab, samp, nullptr, twid, map, 0.f, 0.f, 0, refspec, R, st);
}

int cra_launch_polar_single(const float* image, int nx, const CraRingTab* tab, const CraRingTab& htab,
                            const float2* samp, const float* sampw, const float2* twid, float cx, float cy,
                            int normalize_ring, float* spec, cudaStream_t st)
{
    CraRowMap map{};
    map.row_start = nullptr; map.p0 = 0;
    return launch_polar<0>(image, nx, tab, htab, samp, sampw, twid, map, cx, cy, normalize_ring, spec, 1, st);
}
