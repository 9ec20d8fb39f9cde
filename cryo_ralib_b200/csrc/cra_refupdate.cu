// cra_refupdate.cu -- the per-iteration reference update on the device (the rank-0 section of the reference's
// loop, test_mref.py:238-286; reference-free twin test_reffree.py:314-372):
//
//   class_fsc_kernel    per class: avg = (even + odd) / n (test_mref.py:255-256) into the reference slot, and the
//                       ring-binned Fourier ring correlation sums of fsc(even, odd) (EMData::calc_fourier_shell_
//                       correlation, test_mref.py:254; optionally fsc_mask's in-mask mean subtraction and masking,
//                       test_reffree.py:708): per shell  sum Re(E conj O), sum |E|^2, sum |O|^2  in double.
//   filter_center_kernel  per reference: filt_tanl(avg, fl, aa) -> center_2D(., 1) = phase_cog + fshift(-cs)
//                       (sp_user_functions.ref_ali2d, test_mref.py:273-276) in ONE Fourier round trip: the centre
//                       of gravity of phase_cog is the phase of the (0,1) and (1,0) coefficients of the filtered
//                       spectrum, and the shift is a phase ramp on that same spectrum.  A given shift instead of
//                       the phase centre serves the reference-free driver (fshift(tavg, -cs), test_reffree.py:741-745).
//
// What stays on the host is what the reference keeps in Python as well: the class average of the FSC curves and the
// 2-parameter simplex fit of the tangent filter (cra_fit_tanh) on <= nx/2+1 numbers.
//
// A reference is nx x nx with nx arbitrary (90, 128 in the named configurations), so the transforms are direct
// separable DFTs held in shared memory, as in cra_refavg.cu: ~12 MFLOP per class, microseconds for hundreds of classes.
#include "cra_common.cuh"
#include <math.h>

namespace {

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__device__ __forceinline__ void make_twiddles(float2* tw, int nx, int tid, int nt)
{
    for (int j = tid; j < nx; j += nt) {
        double s, c; sincospi(2.0 * (double)j / (double)nx, &s, &c);
        tw[j] = make_float2((float)c, (float)(-s));          // exp(-2 pi i j / nx)
    }
}

// rows: B[y][kx] = sum_x img[y][x] w^(kx x)
__device__ __forceinline__ void dft_rows(const float* img, float2* B, const float2* tw, int nx, int nh, int tid, int nt)
{
    for (int it = tid; it < nx * nh; it += nt) {
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
}

// one column coefficient: sum_y B[y][kx] w^(ky y)
__device__ __forceinline__ float2 dft_col(const float2* B, const float2* tw, int nx, int nh, int ky, int kx)
{
    float2 acc = make_float2(0.f, 0.f);
    int ph = 0;
    for (int y = 0; y < nx; ++y) {
        const float2 v = cmulf(B[y * nh + kx], tw[ph]);
        acc.x += v.x; acc.y += v.y;
        ph += ky; if (ph >= nx) ph -= nx;
    }
    return acc;
}

// One CTA per class.  shared: A [nx][nh] float2 (first the image), B [nx][nh], E [nx][nh], tw [nx], then the shell sums.
// GS: boxes whose three spectra exceed shared memory (nx beyond ~136) keep A, B, E in a global scratch buffer
// (3 nx nh float2 per class, L2-resident); the arithmetic and its order are the same.
//   shell [nx][nh]: shell index of the half-complex coefficient (ky, kx), or -1 where fsc skips it (host table:
//   Friedel mates on the kx = 0 column, shells beyond nx/2)
//   fsc_out [R][3][nsh] double: num, |E|^2, |O|^2
//   masked: subtract the in-mask mean and multiply by the mask first (fsc_mask)
template <bool GS>
__global__ void __launch_bounds__(256)
class_fsc_kernel(const float* __restrict__ sums, const float* __restrict__ counts, float* __restrict__ refs,
                 const short* __restrict__ shell, const float* __restrict__ mask, int nx, int nsh, int masked, int min_members,
                 int write_avg, float avg_div, double* __restrict__ fsc_out, float2* __restrict__ scratch)
{
    extern __shared__ __align__(16) float2 s_f[];
    const int nh = nx / 2 + 1, nf = nx * nh, npix = nx * nx;
    float2* A = GS ? scratch + (size_t)blockIdx.x * 3 * nf : s_f;
    float2* B = A + nf;
    float2* E = B + nf;
    float2* tw = GS ? s_f : E + nf;
    double* acc = reinterpret_cast<double*>(tw + nx + (nx & 1));        // 3 * nsh doubles (8-byte aligned)
    __shared__ double s_red[2][8];
    float* img = reinterpret_cast<float*>(A);
    const int r = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const float n = counts[r];
    double* out = fsc_out + (size_t)r * 3 * nsh;
    if (n < (float)min_members) {                                        // a vanished class: the host reseeds it
        for (int i = tid; i < 3 * nsh; i += nt) out[i] = 0.0;
        return;
    }
    const float* ev = sums + (size_t)r * 2 * npix;
    const float* od = ev + npix;
    if (write_avg) {
        // avg = (even + odd) * (1 / n)  (Util.add_img, Util.mul_scalar(1.0/float(n)), test_mref.py:255-256)
        const float inv = avg_div > 0.f ? 1.0f / avg_div : 1.0f / n;
        float* dst = refs + (size_t)r * npix;
        for (int i = tid; i < npix; i += nt) dst[i] = (ev[i] + od[i]) * inv;
    }
    make_twiddles(tw, nx, tid, nt);
    for (int i = tid; i < 3 * nsh; i += nt) acc[i] = 0.0;
    for (int half = 0; half < 2; ++half) {
        const float* src = half ? od : ev;
        __syncthreads();
        float mean = 0.f;
        if (masked) {                                                    // in-mask mean in double (Util.infomask)
            double s = 0.0, c = 0.0;
            for (int i = tid; i < npix; i += nt) if (mask[i] > 0.5f) { s += (double)src[i]; c += 1.0; }
            for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
            if ((tid & 31) == 0) { s_red[0][tid >> 5] = s; s_red[1][tid >> 5] = c; }
            __syncthreads();
            double ts = 0.0, tc = 0.0;
            for (int w = 0; w < nt / 32; ++w) { ts += s_red[0][w]; tc += s_red[1][w]; }
            mean = (float)(ts / tc);
            __syncthreads();
        }
        for (int i = tid; i < npix; i += nt) img[i] = masked ? (src[i] - mean) * mask[i] : src[i];
        __syncthreads();
        dft_rows(img, B, tw, nx, nh, tid, nt);
        __syncthreads();
        for (int it = tid; it < nf; it += nt) {
            const int ky = it / nh, kx = it - ky * nh;
            const float2 v = dft_col(B, tw, nx, nh, ky, kx);
            if (half == 0) { E[it] = v; continue; }
            const int sh = shell[it];
            if (sh < 0) continue;
            const float2 e = E[it];
            atomicAdd(&acc[sh], (double)e.x * (double)v.x + (double)e.y * (double)v.y);
            atomicAdd(&acc[nsh + sh], (double)e.x * (double)e.x + (double)e.y * (double)e.y);
            atomicAdd(&acc[2 * nsh + sh], (double)v.x * (double)v.x + (double)v.y * (double)v.y);
        }
    }
    __syncthreads();
    for (int i = tid; i < 3 * nsh; i += nt) out[i] = acc[i];
}

// One CTA per reference, in place.  mode 0: filter only; 1: filter, then centre by the phase centre of gravity;
// 2: filter, then shift by (-shift.x, -shift.y) (the same shift for every image).  cs_out [n][2]: the shift removed.
template <bool GS>
__global__ void __launch_bounds__(256)
filter_center_kernel(float* __restrict__ imgs, int nx, float fl, float aa, int mode, float2 shift, float* __restrict__ cs_out,
                     float2* __restrict__ scratch)
{
    extern __shared__ __align__(16) float2 s_f[];
    const int nh = nx / 2 + 1, nf = nx * nh;
    float2* A = GS ? scratch + (size_t)blockIdx.x * 2 * nf : s_f;
    float2* B = A + nf;
    float2* tw = GS ? s_f : B + nf;
    __shared__ float s_cs[2];
    float* img = reinterpret_cast<float*>(A);
    float* g = imgs + (size_t)blockIdx.x * nx * nx;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < nx * nx; i += nt) img[i] = g[i];
    make_twiddles(tw, nx, tid, nt);
    __syncthreads();
    dft_rows(img, B, tw, nx, nh, tid, nt);
    __syncthreads();
    // columns + filter: A[ky][kx] = H(ky, kx) sum_y B[y][kx] w^(ky y)
    const float c = 3.14159265358979323846f / (2.0f * aa * fl);
    for (int it = tid; it < nf; it += nt) {
        const int ky = it / nh, kx = it - ky * nh;
        const float2 v = dft_col(B, tw, nx, nh, ky, kx);
        const float fy = (float)((ky > nx / 2) ? ky - nx : ky) / (float)nx, fx = (float)kx / (float)nx;
        const float d = sqrtf(fx * fx + fy * fy);
        const float H = 0.5f * (tanhf(c * (d + fl)) - tanhf(c * (d - fl)));
        A[it] = make_float2(v.x * H, v.y * H);
    }
    __syncthreads();
    if (mode != 0) {
        if (tid == 0) {
            float sx = shift.x, sy = shift.y;
            if (mode == 1) {
                // EMData::phase_cog: marginal over y as a function of x has the k = 1 coefficient F(ky=0, kx=1);
                // f1 = atan2(sum sin . marg, sum cos . marg) = atan2(-Im F, Re F); cs = f1 / (2 pi / n) + 1 - (n/2 + 1)
                const double twopi = 6.283185307179586476925;
                double f1 = atan2(-(double)A[1].y, (double)A[1].x);
                if (f1 < 0.0) f1 += twopi;
                sx = (float)(f1 / (twopi / nx) + 1.0 - (double)(nx / 2 + 1));
                f1 = atan2(-(double)A[nh].y, (double)A[nh].x);
                if (f1 < 0.0) f1 += twopi;
                sy = (float)(f1 / (twopi / nx) + 1.0 - (double)(nx / 2 + 1));
            }
            s_cs[0] = sx; s_cs[1] = sy;
            if (cs_out) { cs_out[2 * blockIdx.x] = sx; cs_out[2 * blockIdx.x + 1] = sy; }
        }
        __syncthreads();
        // fshift(img, -cs): multiply by exp(+2 pi i (fx cs_x + fy cs_y))
        const float sx = s_cs[0], sy = s_cs[1];
        for (int it = tid; it < nf; it += nt) {
            const int ky = it / nh, kx = it - ky * nh;
            const float fy = (float)((ky > nx / 2) ? ky - nx : ky) / (float)nx, fx = (float)kx / (float)nx;
            float s, co; sincospif(2.0f * (fx * sx + fy * sy), &s, &co);
            A[it] = cmulf(A[it], make_float2(co, s));
        }
        __syncthreads();
    } else if (cs_out && tid == 0) { cs_out[2 * blockIdx.x] = 0.f; cs_out[2 * blockIdx.x + 1] = 0.f; }
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
    // inverse rows of a Hermitian spectrum.  After a phase ramp the Nyquist column of an even box is no longer
    // real; numpy's irfft2 (the host twin) takes its real part, and so does the sum below.
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

size_t cra_class_fsc_smem(int nx, int nsh)
{
    const size_t nh = (size_t)nx / 2 + 1;
    return (3 * (size_t)nx * nh + nx + (nx & 1)) * sizeof(float2) + 3 * (size_t)nsh * sizeof(double);
}

// float2 elements of global scratch one image needs when its spectra do not fit into shared memory (0: they fit)
size_t cra_dft_scratch_elems(int nx, int nsh, int nspec)
{
    const size_t nh = (size_t)nx / 2 + 1;
    const size_t smem = ((size_t)nspec * nx * nh + nx + (nx & 1)) * sizeof(float2) + 3 * (size_t)nsh * sizeof(double);
    int dev = 0, lim = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
    return (smem + 1024 > (size_t)lim) ? (size_t)nspec * nx * nh : 0;
}

int cra_launch_class_fsc(const float* sums, const float* counts, float* refs, const short* shell, const float* mask, int R,
                         int nx, int nsh, int masked, int min_members, int write_avg, float avg_div, double* fsc_out,
                         float2* scratch, cudaStream_t st)
{
    if (R <= 0) return 0;
    if (cra_dft_scratch_elems(nx, nsh, 3)) {
        if (!scratch) { cra_set_error("device reference update: no scratch buffer for a box this large"); return 1; }
        const size_t smem = ((size_t)nx + (nx & 1)) * sizeof(float2) + 3 * (size_t)nsh * sizeof(double);
        class_fsc_kernel<true><<<R, 256, smem, st>>>(sums, counts, refs, shell, mask, nx, nsh, masked, min_members, write_avg,
                                                     avg_div, fsc_out, scratch);
    } else {
        const size_t smem = cra_class_fsc_smem(nx, nsh);
        if (cra_ensure_dyn_smem(reinterpret_cast<const void*>(&class_fsc_kernel<false>), smem)) return 1;
        class_fsc_kernel<false><<<R, 256, smem, st>>>(sums, counts, refs, shell, mask, nx, nsh, masked, min_members, write_avg,
                                                      avg_div, fsc_out, nullptr);
    }
    CRA_CUDA(cudaGetLastError());
    return 0;
}

int cra_launch_filter_center(float* imgs, int n, int nx, float fl, float aa, int mode, float sx, float sy, float* cs_out,
                             float2* scratch, cudaStream_t st)
{
    if (n <= 0) return 0;
    if (!(fl > 0.f) || !(aa > 0.f)) { cra_set_error("tangent filter: cut-off and fall-off must be positive"); return 1; }
    const size_t nh = (size_t)nx / 2 + 1;
    if (cra_dft_scratch_elems(nx, 0, 2)) {
        if (!scratch) { cra_set_error("tangent filter: no scratch buffer for a box this large"); return 1; }
        filter_center_kernel<true><<<n, 256, (size_t)nx * sizeof(float2), st>>>(imgs, nx, fl, aa, mode, make_float2(sx, sy), cs_out, scratch);
    } else {
        const size_t smem = (2 * (size_t)nx * nh + nx) * sizeof(float2);
        if (cra_ensure_dyn_smem(reinterpret_cast<const void*>(&filter_center_kernel<false>), smem)) return 1;
        filter_center_kernel<false><<<n, 256, smem, st>>>(imgs, nx, fl, aa, mode, make_float2(sx, sy), cs_out, nullptr);
    }
    CRA_CUDA(cudaGetLastError());
    return 0;
}
