// cra_refavg.cu -- references rebuilt on the device from the class sums, and the tangent low-pass
// applied to them: the reference-side half of gpu_isac's class-bound reference-free alignment
// (BatchHandler::fetch_averages, cuda/gpu_aln_noref.cu:743-775; ref_free_alignment_2D_filter_references
// -> cu_apply_tanl_filter_to_tex, cuda/gpu_aln_noref.cu:777-814).  The filter has EMAN2 filt_tanl
// semantics (TANH_LOW_PASS on the un-padded nx x nx transform, H = 0.5 (tanh(c (d + fl)) - tanh(c (d - fl))),
// c = pi / (2 aa fl), d = |frequency| in cycles per pixel), which is also the formula the reference's
// kernel evaluates.
//
// A reference is nx x nx with nx arbitrary (90 in the named configs), so the transform is a direct
// separable DFT held in shared memory (filter_center_kernel, cra_refupdate.cu) -- 4 passes of nx^2 (nx/2+1)
// complex MACs per image, microseconds for the few hundred references of an iteration; no cuFFT plan.
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

}  // namespace

int cra_launch_class_average(const float* sums, const float* counts, float* refs, int R, int nx, cudaStream_t st)
{
    if (R <= 0) return 0;
    class_average_kernel<<<R, 256, 0, st>>>(sums, counts, refs, nx * nx);
    CRA_CUDA(cudaGetLastError());
    return 0;
}

// filt_tanl in place: the filter-only mode of filter_center_kernel (cra_refupdate.cu), one Fourier round trip per image
int cra_launch_tanl_filter(float* imgs, int n, int nx, float fl, float aa, float2* scratch, cudaStream_t st)
{
    return cra_launch_filter_center(imgs, n, nx, fl, aa, 0, 0.f, 0.f, nullptr, scratch, st);
}
