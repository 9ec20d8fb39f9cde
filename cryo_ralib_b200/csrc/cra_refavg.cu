// cra_refavg.cu -- references rebuilt on the device from the class sums, and the tangent low-pass
// applied to them: the reference-side half of gpu_isac's class-bound reference-free alignment
// (BatchHandler::fetch_averages, cuda/gpu_aln_noref.cu:743-775; ref_free_alignment_2D_filter_references
// -> cu_apply_tanl_filter_to_tex, cuda/gpu_aln_noref.cu:777-814).  The filter has EMAN2 filt_tanl
// semantics (TANH_LOW_PASS on the un-padded nx x nx transform, H = 0.5 (tanh(c (d + fl)) - tanh(c (d - fl))),
// c = pi / (2 aa fl), d = |frequency| in cycles per pixel), which is also the formula the reference's
// kernel evaluates.
//
// A reference is nx x nx with nx arbitrary (90 in the named configs), so the transform is a direct
// separable DFT held entirely in shared memory -- 4 passes of nx^2 (nx/2+1) complex MACs per image,
// microseconds for the few hundred references of an iteration; no cuFFT plan, no HBM round trip.
#include "cra_common.cuh"
#include <math.h>

namespace {

// refs[r] = (even[r] + odd[r]) / count[r]; classes without members keep their reference
__global__ void __launch_bounds__(256)
class_average_kernel(const float* __restrict__ sums, const float* __restrict__ counts, float* __restrict__ refs, int npix)
{
    const int r = blockIdx.x;
    const float n = counts[r];
    if (!(n > 0.5f)) return;
    const float* e = sums + (size_t)r * 2 * npix;
    const float* o = e + npix;
    float* dst = refs + (size_t)r * npix;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) dst[i] = (e[i] + o[i]) / n;
}

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// one CTA per image; shared memory: A [nx][nh] float2 (first the image itself), B [nx][nh] float2, tw [nx] float2
__global__ void __launch_bounds__(256)
tanl_filter_kernel(float* __restrict__ imgs, int nx, float fl, float aa)
{
    extern __shared__ __align__(16) float2 s_f[];
    const int nh = nx / 2 + 1, nf = nx * nh;
    float2* A = s_f;
    float2* B = s_f + nf;
    float2* tw = B + nf;                                  // tw[j] = exp(-2 pi i j / nx)
    float* img = reinterpret_cast<float*>(A);
    float* g = imgs + (size_t)blockIdx.x * nx * nx;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < nx * nx; i += nt) img[i] = g[i];
    for (int j = tid; j < nx; j += nt) {
        double s, c; sincospi(2.0 * (double)j / (double)nx, &s, &c);
        tw[j] = make_float2((float)c, (float)(-s));
    }
    __syncthreads();
    // rows: B[y][kx] = sum_x img[y][x] w^(kx x)
    for (int it = tid; it < nf; it += nt) {
        const int y = it / nh, kx = it - y * nh;
        const float* row = img + y * nx;
        float2 acc = make_float2(0.f, 0.f);
        int ph = 0;
        for (int x = 0; x < nx; ++x) {
            const float2 w = tw[ph];
            acc.x = fmaf(row[x], w.x, acc.x); acc.y = fmaf(row[x], w.y, acc.y);
            ph += kx; if (ph >= nx) ph -= nx;
        }
        B[it] = acc;
    }
    __syncthreads();
    // columns + filter: A[ky][kx] = H(ky, kx) sum_y B[y][kx] w^(ky y)
    const float c = 3.14159265358979323846f / (2.0f * aa * fl);
    for (int it = tid; it < nf; it += nt) {
        const int ky = it / nh, kx = it - ky * nh;
        float2 acc = make_float2(0.f, 0.f);
        int ph = 0;
        for (int y = 0; y < nx; ++y) {
            const float2 v = cmulf(B[y * nh + kx], tw[ph]);
            acc.x += v.x; acc.y += v.y;
            ph += ky; if (ph >= nx) ph -= nx;
        }
        const float fy = (float)((ky > nx / 2) ? ky - nx : ky) / (float)nx, fx = (float)kx / (float)nx;
        const float d = sqrtf(fx * fx + fy * fy);
        const float H = 0.5f * (tanhf(c * (d + fl)) - tanhf(c * (d - fl)));
        A[it] = make_float2(acc.x * H, acc.y * H);
    }
    __syncthreads();
    // inverse columns: B[y][kx] = sum_ky A[ky][kx] conj(w)^(ky y)
    for (int it = tid; it < nf; it += nt) {
        const int y = it / nh, kx = it - y * nh;
        float2 acc = make_float2(0.f, 0.f);
        int ph = 0;
        for (int ky = 0; ky < nx; ++ky) {
            const float2 w = tw[ph];
            const float2 v = cmulf(A[ky * nh + kx], make_float2(w.x, -w.y));
            acc.x += v.x; acc.y += v.y;
            ph += y; if (ph >= nx) ph -= nx;
        }
        B[it] = acc;
    }
    __syncthreads();
    // inverse rows of a Hermitian spectrum: out[y][x] = (Re B[y][0] + sum_kx>0 m Re(B[y][kx] conj(w)^(kx x))) / nx^2
    const float inv = 1.0f / ((float)nx * (float)nx);
    const bool even = (nx & 1) == 0;
    for (int it = tid; it < nx * nx; it += nt) {
        const int y = it / nx, x = it - y * nx;
        const float2* row = B + y * nh;
        float acc = row[0].x;
        int ph = 0;
        for (int kx = 1; kx < nh; ++kx) {
            ph += x; if (ph >= nx) ph -= nx;
            const float2 w = tw[ph];
            const float re = row[kx].x * w.x + row[kx].y * w.y;          // Re(B conj(w))
            acc += ((even && kx == nh - 1) ? 1.0f : 2.0f) * re;
        }
        g[it] = acc * inv;
    }
}

}  // namespace

int cra_launch_class_average(const float* sums, const float* counts, float* refs, int R, int nx, cudaStream_t st)
{
    if (R <= 0) return 0;
    class_average_kernel<<<R, 256, 0, st>>>(sums, counts, refs, nx * nx);
    CRA_CUDA(cudaGetLastError());
    return 0;
}

int cra_launch_tanl_filter(float* imgs, int n, int nx, float fl, float aa, cudaStream_t st)
{
    if (n <= 0) return 0;
    if (!(fl > 0.f) || !(aa > 0.f)) { cra_set_error("tangent filter: cut-off and fall-off must be positive"); return 1; }
    const size_t nh = (size_t)nx / 2 + 1;
    const size_t smem = (2 * (size_t)nx * nh + nx) * sizeof(float2);
    int dev = 0, lim = 0;
    CRA_CUDA(cudaGetDevice(&dev));
    CRA_CUDA(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem > (size_t)lim) { cra_set_error("tangent filter: image too large for the shared-memory transform"); return 1; }
    if (cra_ensure_dyn_smem(reinterpret_cast<const void*>(&tanl_filter_kernel), smem)) return 1;
    tanl_filter_kernel<<<n, 256, smem, st>>>(imgs, nx, fl, aa);
    CRA_CUDA(cudaGetLastError());
    return 0;
}
