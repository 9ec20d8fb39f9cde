// cra_rotsum.cu -- stage 5 of the hot path: rot_shift2D (EMAN2
// rot_scale_trans2D_background + quadri_background + xform.mirror x) fused with the
// even/odd class-sum accumulation of test_mref.py:210-215 (Util.add_img into
// refi[iref][im%2], count += 1).  Replaces cu_transform_batch + cu_average_batch_m /
// the CuPy kernel_sum_oe (cuda/gpu_aln_noref.cu:1145-1197, :1232-1274;
// test_mref_gpu_align.py:48-80).  The in-repo restatement of the interpolation
// this follows is notebook/02_CuPy_Image_Processing_rot_shift2d.ipynb cell 2.
//
// One CTA per particle: the image is staged in shared memory, every thread produces
// output pixels and adds them straight into the L2-resident [R][2][nx][nx] sums with
// float atomics (RED), so the transformed image never goes to HBM unless asked for.
// Also holds the FP32 FMA peak micro-benchmark used as the roofline denominator of
// the CCF kernel.
#include "cra_common.cuh"
#include "cra_tma.cuh"

namespace {

__device__ __forceinline__ float quadri_bg(float x, float y, int nx, const float* __restrict__ f, int xnew, int ynew)
{
    if ((x < 1.0f) || (x >= (float)(nx + 1)) || (y < 1.0f) || (y >= (float)(nx + 1))) {
        x = (float)xnew; y = (float)ynew;
    }
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
    float c5 = f[(jp1 - 1) * nx - 1 + ip1] - f0 - c1 - c3;
    return f0 + dx0 * (c1 + (dx0 - 1.0f) * c2 + dy0 * c5) + dy0 * (c3 + (dy0 - 1.0f) * c4);
}

__device__ __forceinline__ float restrict2(float x, int nx)
{
    while (x >= (float)nx) x -= nx;
    while (x <= -(float)nx) x += nx;
    return x;
}

// GIMG: the image does not fit into shared memory (boxes beyond ~238 pixels): the taps come from global memory
template <bool GIMG>
__global__ void __launch_bounds__(256)
rotsum_kernel(const float* __restrict__ images, int nx, int p0, const float4* __restrict__ params,
              const int* __restrict__ iref, long global_offset, float* __restrict__ sums,
              float* __restrict__ counts, float* __restrict__ out_images)
{
    extern __shared__ __align__(16) float s_img[];
    const int p = blockIdx.x;
    const int ref = iref ? iref[p] : 0;
    if (iref && ref < 0) return;
    const int npix = nx * nx;
    const float* img = images + (size_t)(p0 + p) * npix;
    // the image tile: one bulk asynchronous copy (TMA) when the image is 16-byte granular, else a cooperative load
    __shared__ __align__(8) unsigned long long s_bar;
    const bool bulk = !GIMG && cratma::bulk_ok(img, (size_t)npix * sizeof(float));
    if (GIMG) {
    } else if (bulk) {
        if (threadIdx.x == 0) {
            cratma::mbar_init(&s_bar, 1);
            cratma::bulk_load(s_img, img, (unsigned)(npix * sizeof(float)), &s_bar);
        }
    } else {
        for (int i = threadIdx.x; i < npix; i += blockDim.x) s_img[i] = __ldg(img + i);
    }
    const float4 pr = params[p];
    __syncthreads();                           // the barrier's initialisation (or the cooperative load) is visible
    if (bulk) cratma::mbar_wait(&s_bar, 0);
    const float ang = pr.x;                    // radians, rounded to float by the host (cra_api.cu: pack_par)
    const float delx = restrict2(pr.y, nx), dely = restrict2(pr.z, nx);
    const int mirror = pr.w > 0.5f;
    const int xc = nx / 2, yc = nx / 2;
    const float shiftxc = xc + delx, shiftyc = yc + dely;
    // Source coordinates exactly as the host library forms them: a correctly rounded float cosine / sine (the
    // device's cosf is 1-2 ulp) and separately rounded products and sums (no FMA contraction).  quadri_background is
    // discontinuous across cell borders (its c2 / c4 terms belong to the cell), so a last-ulp difference of xold on a
    // pixel boundary would move that output pixel by O(1) -- seen on 5 % of the particles before this.
    const float cang = (float)cos((double)ang), sang = (float)sin((double)ang);
    const int x_start = 1 - nx % 2;
    const int parity = (int)((global_offset + p) & 1);
    float* dst = sums ? sums + ((size_t)ref * 2 + parity) * npix : nullptr;
    float* oimg = out_images ? out_images + (size_t)p * npix : nullptr;
    for (int idx = threadIdx.x; idx < npix; idx += blockDim.x) {
        const int iy = idx / nx, ix = idx - iy * nx;
        const float y = __fsub_rn((float)iy, shiftyc);
        const float ycang = __fadd_rn(__fmul_rn(y, cang), (float)yc);
        const float ysang = __fadd_rn(__fmul_rn(-y, sang), (float)xc);
        const float x = __fsub_rn((float)ix, shiftxc);
        const float xold = __fadd_rn(__fmul_rn(x, cang), ysang);
        const float yold = __fadd_rn(__fmul_rn(x, sang), ycang);
        const float v = GIMG ? quadri_bg(__fadd_rn(xold, 1.0f), __fadd_rn(yold, 1.0f), nx, img, ix + 1, iy + 1)
                             : quadri_bg(__fadd_rn(xold, 1.0f), __fadd_rn(yold, 1.0f), nx, s_img, ix + 1, iy + 1);
        const int ixd = (mirror && ix >= x_start) ? (x_start + nx - 1 - ix) : ix;
        const int o = iy * nx + ixd;
        if (dst) atomicAdd(dst + o, v);
        if (oimg) oimg[o] = v;
    }
    if (counts && threadIdx.x == 0) atomicAdd(counts + ref, 1.0f);
}

// ---------------------------------------------------------------- FP32 peak
template <int PACKED>
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float seed)
{
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) a[i] = __ffma2_rn(a[i], m, c);
            else { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    if (s == 123.456f) out[0] = s;
}

template <int PACKED>
int run_peak(double* tf)
{
    float* d = nullptr;
    CRA_CUDA(cudaMalloc(&d, 4));
    int dev = 0, sms = 0;
    CRA_CUDA(cudaGetDevice(&dev));
    CRA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int iters = 8192, grid = sms * 8;
    cudaEvent_t e0, e1;
    CRA_CUDA(cudaEventCreate(&e0)); CRA_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CRA_CUDA(cudaEventRecord(e0));
        fma_peak_kernel<PACKED><<<grid, 256>>>(d, iters, 1.0f + rep);
        CRA_CUDA(cudaEventRecord(e1));
        CRA_CUDA(cudaEventSynchronize(e1));
        float ms = 0; CRA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 16.0 * iters * 256.0 * grid;
        double t = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && t > best) best = t;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tf = best;
    return 0;
}

}  // namespace

int cra_launch_rotsum(const float* images, int nx, int p0, int n, const float4* params, const int* iref,
                      long global_offset, float* sums, float* counts, float* out_images, cudaStream_t st)
{
    if (n <= 0) return 0;
    const size_t smem = (((size_t)nx * nx + 3) & ~(size_t)3) * sizeof(float);
    int dev = 0, lim = 0;
    CRA_CUDA(cudaGetDevice(&dev));
    CRA_CUDA(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem + 1024 > (size_t)lim) {
        rotsum_kernel<true><<<n, 256, 0, st>>>(images, nx, p0, params, iref, global_offset, sums, counts, out_images);
    } else {
        if (cra_ensure_dyn_smem(reinterpret_cast<const void*>(&rotsum_kernel<false>), smem)) return 1;
        rotsum_kernel<false><<<n, 256, smem, st>>>(images, nx, p0, params, iref, global_offset, sums, counts, out_images);
    }
    CRA_CUDA(cudaGetLastError());
    return 0;
}

int cra_fp32_peak(double* tf_ffma, double* tf_ffma2)
{
    if (run_peak<0>(tf_ffma)) return 1;
    if (run_peak<1>(tf_ffma2)) return 1;
    return 0;
}
