// cra_common.cuh -- shared declarations of the sm_100a alignment engine.
//
// Data layout in HBM (all float32 unless noted):
//   images   [P][nx][nx]            resident particle stack (mask-mean subtracted)
//   refs     [R][nx][nx]            current references
//   refspec  [R/4][nc][4] float2    weighted reference spectra (Applyws applied), same layout
//   spec     [rows/4][2*nc] float4  particle spectra of one row batch; a row is one (particle,
//                                   shift) pair in the reference's visit order.  Four consecutive
//                                   rows form a group; per ring the group stores two planes of
//                                   len/2+1 float4, plane 0 = rows (0,1), plane 1 = rows (2,3), so
//                                   the contraction fetches 4 rows with two fully coalesced
//                                   128-bit loads (see "device spectrum")
//   cand     [rows][ntile_n]        best (value, code) of each row x reference tile
//   sums     [R][2][nx][nx] + [R]   even/odd class sums followed by counts
// Host-visible spectra use the SPIDER packed per-ring layout of Util.Frngs: ring i occupies
// floats [off_i, off_i+len_i): slot0 = Re F_0, slot1 = Re F_{len/2}, then
// (Re,Im) F_k for k = 1..len/2-1, F_k = sum_n x_n exp(-2 pi i n k / len).
// The DEVICE spectrum un-packs that: ring i holds len_i/2+1 complex values F_0..F_{len/2}
// (imaginary parts of F_0 and F_{len/2} are 0) at complex offset coff_i; nc = lcirc/2 + nring.
// With it Crosrng_ms needs no special cases: every (ring, k <= len/2) is one complex MAC, and
// the Nyquist term of a short ring lands on frequency len/2 exactly as q(numr3i+1) does.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string>
#include <vector>
#include "../../include/cryo_ralib.h"
#include <nvtx3/nvToolsExt.h>

#define CRA_MAX_RINGS 192

struct CraRingTab {            // device-resident ring table (Numrinit triplets, 0-based offsets)
    int   nring, lcirc, maxrin, log2n;
    int   off[CRA_MAX_RINGS];
    int   len[CRA_MAX_RINGS];
    int   rad[CRA_MAX_RINGS];
    int   coff[CRA_MAX_RINGS];   // complex offset of ring i in the device spectrum
    int   nc;                    // complex elements per device spectrum = lcirc/2 + nring
    int   poff[CRA_MAX_RINGS];   // float2 offset of ring i in the polar kernel's padded smem buffer
    int   lcpad;                 // floats per padded smem ring buffer (rows of NB complex + 1 pad)
    float wr[CRA_MAX_RINGS];     // ringwe (Applyws)
    float wn[CRA_MAX_RINGS];     // Normalize_ring weight r*2pi/len
    float nn;                    // sum of wn over every sample, accumulated in float like Normalize_ring
};

#ifndef CRA_POLAR_RPB
#define CRA_POLAR_RPB 2        // shift rows resampled per CTA of the polar kernel (1, 2 or 4)
#endif

struct CraPolarItems {         // flat work lists of the ring FFT passes (device pointers)
    const int* A; int nA;      // (ring << 16 | column b)   pass A: NA-point DFTs
    const int* B; int nB;      // (ring << 16 | row ka)     pass B: NB-point DFTs
    const int* C; int nC;      // (ring << 16 | k)          pass C: split + store, k <= len/4
    const int* Cg; int nCg;    // same without the k = 0 items (grouped row kernel), per phase in an order whose
                               // half-warps hit 16 distinct shared-memory banks (build_group_plan)
    const int* D; int nD;      // (unit << 16 | k)          pass D of the grouped kernel: unit gather, same ordering rule
};

// ---- grouped row kernel (cra_polar_grp.cu) -----------------------------------------------------
#ifndef CRA_GRP_RMAX
#define CRA_GRP_RMAX 17        // most shift rows one CTA of the grouped row kernel resamples together
#endif
struct CraPhase {              // one walk of the CTA over a set of consecutive 4-ring units
    int q0, q1;                // quarter-ring sample range in samp[]
    int a0, a1, b0, b1, c0, c1;// item ranges in CraPolarItems A / B / Cg
    int d0;                    // first item of the phase in CraPolarItems D (upr items)
    int u0, u1;                // ring units [u0, u1): unit u = slots 4u..4u+3, slot s <-> ring nring-1-s
    int upr;                   // pass-D lanes per row = sum over the units of their longest half length
    int magicA, magicB, magicC, magicD;   // floor(2^24 / n) + 1 for n = items A, B, C, upr (fastdiv)
    // row sets of passes C / D for a block of nr rows: lanes = n * nset, each lane loops ceil(nr/nset) rows;
    // the value that minimises ceil(n*nset/256) * (lane set-up + ceil(nr/nset) * per-row cost)
    unsigned char nsetC[CRA_GRP_RMAX + 1], nsetD[CRA_GRP_RMAX + 1];
};
struct CraGroupPlan {
    const CraPhase* phases;    // device
    int nphase;
    const int* ppoff;          // [nring] float2 offset of ring i inside its phase's row buffer (device)
    const int* unit_nk;        // [units] longest half length (len/2 of its first slot) of unit u (device)
    int stride;                // floats per row of the phase buffer
    int rmax;                  // rows per CTA (<= CRA_GRP_RMAX), chosen for shared-memory fit
    int nh;                    // independent thread groups of a CTA: 2 when a group still has >= 6 rows, else 1
    int nring;
    int tile;                  // 0: the shared-memory tile is the whole image; else the side (a multiple of 4) of a square
                               // window around the particle's search window (boxes too large for, or much larger than, the ring set)
};

// Phase classes per axis of the grouped row kernel: 1 for a whole-pixel step, 2 or 4 for a step of 1/2 or 1/4
// pixel, 0 when the step has no such structure (the general per-row kernel serves it).
__host__ __device__ __forceinline__ int cra_group_sub(float step)
{
    if (step >= 1.0f) return (step == floorf(step) && step < 1024.f) ? 1 : 0;
    if (step == 0.5f) return 2;
    if (step == 0.25f) return 4;
    return 0;
}

struct CraRowMap {             // how rows of the current batch map to particles
    const int*       row_start;  // [np+1] first row of each batch-local particle
    const int*       chunk_start;// [np+1] first polar CTA of each batch-local particle
    int              nchunks;
    const CraSearch* search;     // [np]   per-particle request (batch-local)
    const int4*      win;        // [np]   lkx, rkx, lky, rky
    int              np;
    int              nrows;
    int              p0;         // slot of batch-local particle 0 in the resident stack
    const float*     dc;         // [max_particles] in-mask mean each resident image still carries (may be null)
    float            step;
};

// float2 index, inside a 4-row group, of (row r, ring with complex offset coff and len/2 = half, k)
__host__ __device__ __forceinline__ size_t cra_spec_idx(int coff, int half, int r, int k)
{
    return ((size_t)2 * coff + (size_t)((r & 3) >> 1) * (half + 1) + k) * 2 + (r & 1);
}

// ---- fragment spectrum ("F16" format, CRA_FMT_FRAG) --------------------------------------------
// Operand layout of the tensor-core contraction (cra_ccf_mma.cu).  For every angular frequency
// k <= maxrin/2 the rings that reach it (len/2 >= k; K_k of them, longest first: slot s <-> ring
// nring-1-s) are cut into chunks of 16 slots (zero padded).  One chunk of one row (or reference)
// is 128 bytes: 4 quads t x [re unit | im unit]; a 16-byte unit holds the 4 slots 16c+4t .. +3 of
// that part as bf16 hi x4 then bf16 lo x4, value = hi + lo (split-bf16: ~17 significant bits).
// That is the REFERENCE layout (B operand: one 128-bit load per thread).  PARTICLE ROWS use the A-operand
// order instead: the 32 bytes of quad t are 8 words  hi{re s0s1, im s0s1, re s2s3, im s2s3} then
// lo{same}  (s0..s3 = slots 16c+4t..+3, first slot in the low half-word), so that one 256-bit load
// is a thread's A fragment with the hi and lo registers already in mma operand order.
// Chunks are stored k-major: chunk index koff[k] + c; a row is (nch + 1) * 128 bytes: chunk nch is never
// written and stays ZERO (the buffers are cleared at allocation) -- the contraction pads its per-warp
// work lists to whole groups with it, so the inner loop has no tail conditionals.
//
// Deferred Normalize_ring: the row kernels may store the spectrum of the RAW polar image and put
// (avg, 1/sigma) of Normalize_ring into norm[row].  (x - avg)/sigma only changes the DC bin of each
// ring by -avg*len, so every lag of a (row, ref) correlation moves by the same -avg * tref[ref],
// tref[ref] = sum_rings len * (weighted reference DC); the CCF kernel applies
// value = (raw - avg * tref) / sigma when it emits a candidate.  norm = (0, 1) means "already normalised".
#define CRA_FMT_F32  0          // planar-pair float2 layout above (FP32 FMA contraction)
#define CRA_FMT_FRAG 1          // fragment layout (tensor-core contraction)
struct CraFragTab {
    int nk;                      // maxrin/2 + 1
    int nch;                     // chunks per row
    const int* koff;             // [nk+1] first chunk of frequency k (device)
    const int* chunk_k;          // [nch]  k << 4 | c (device)
    int unit_rows;               // 1: particle rows use the reference layout [re unit | im unit] (tcgen05 kernel,
                                 // a unit = one UMMA core-matrix row); 0: the mma.sync A-operand order above
};
__host__ __device__ __forceinline__ size_t cra_frag_row_bytes(int nch) { return (size_t)(nch + 1) * 128; }

struct CraCand { float v; int code; };   // code = iref*8192 + mirror*4096 + j   (j = 1-based lag index)

// NVTX range around a host-side phase (the reference marks the same phases from Python with
// cupy.cuda.nvtx.RangePush/Pop, test_mref_gpu_align.py:89, :329, :416, :448); free when no tool is attached
struct CraNvtx { explicit CraNvtx(const char* name) { nvtxRangePushA(name); } ~CraNvtx() { nvtxRangePop(); } };

int cra_host_threads();       // team size of the host-side OpenMP loops (cra_host.cu)

// error plumbing -------------------------------------------------------------
void cra_set_error(const std::string& msg);
#define CRA_CUDA(call)                                                               \
    do { cudaError_t e__ = (call); if (e__ != cudaSuccess) {                         \
        cra_set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " @" +   \
                      __FILE__ + ":" + std::to_string(__LINE__)); return 1; } } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE function attribute: raise it for `func` on the
// current device when `bytes` exceeds what was set there before (thread-safe; cra_api.cu)
int cra_ensure_dyn_smem(const void* func, size_t bytes);

// launchers (each defined in its own .cu; all asynchronous on `st`) ----------
int cra_launch_mask_normalize(float* imgs, int n, int nx, const float* mask, int mode, float* dc_out, cudaStream_t st);
// particle rows described by map -> spec[row]; references -> refspec (weights applied)
// twid_fwd[j] = exp(-2 pi i j / maxrin), j < maxrin
// the general row kernel's variant for a geometry: rows per CTA (want_rpb or 1) and whether the image tile fits
// into shared memory beside the polar rows (gimg = 1: taps read from global memory); 1 = not even one polar row fits
int cra_polar_default_rpb();
int cra_polar_general_layout(int nx, const CraRingTab& htab, size_t smem_limit, int want_rpb, int* rpb, int* gimg);
// norm: [rows] (avg, 1/sigma) written by the row kernels (FRAG format only, may be null otherwise);
// tref: [R] written by the reference kernel (FRAG format only)
int cra_launch_polar_rows(const float* images, int nx, const CraRingTab* tab, const CraRingTab& htab,
                          const float4* samp, const float* sampw, const float2* twid_fwd, const CraPolarItems& items,
                          CraRowMap map, int normalize_ring, float* spec, int fmt, const CraFragTab& frag,
                          float2* norm, int rpb, int gimg, cudaStream_t st);
int cra_launch_polar_refs(const float* refs, int R, int nx, const CraRingTab* tab, const CraRingTab& htab,
                          const float4* samp, const float2* twid_fwd, const CraPolarItems& items, float* refspec,
                          int fmt, const CraFragTab& frag, float* tref, int gimg, cudaStream_t st);
int cra_launch_polar_single(const float* image, int nx, const CraRingTab* tab, const CraRingTab& htab,
                            const float4* samp, const float* sampw, const float2* twid_fwd, const CraPolarItems& items,
                            float cx, float cy, int normalize_ring, float* spec, int fmt, const CraFragTab& frag,
                            float2* norm, int gimg, cudaStream_t st);
size_t cra_polar_group_smem(int nx, int maxrin, const CraGroupPlan& plan);
int cra_launch_polar_group(const float* images, int nx, const CraRingTab* tab, const CraRingTab& htab,
                           const float4* samp, const float2* twid_fwd, const CraPolarItems& items, const CraGroupPlan& plan,
                           CraRowMap map, int normalize_ring, float* spec, const CraFragTab& frag, float2* norm, cudaStream_t st);
int cra_launch_ccf_mma(const unsigned char* spec, int nrows, const unsigned char* refspec, int R, const CraRingTab& htab,
                       const CraFragTab& frag, const std::vector<int>& h_koff, const float2* twid, CraCand* cand,
                       int ntile_n, const float2* norm, const float* tref, cudaStream_t st);
int cra_ccf_mma_num_tiles(int R, int log2n);
// the same contraction with W staged in tensor memory, two CTAs per SM (cra_ccf_tm.cu)
bool cra_ccf_tm_supported(int log2n);
int cra_ccf_tm_num_tiles(int R, int log2n);
// *sched: the context's work lists (built on the first launch, released with cra_ccf_tm_sched_free)
int cra_launch_ccf_tm(const unsigned char* spec, int nrows, const unsigned char* refspec, int R, const CraRingTab& htab,
                      const CraFragTab& frag, const std::vector<int>& h_koff, const float2* twid, CraCand* cand,
                      int ntile_n, const float2* norm, const float* tref, void** sched, cudaStream_t st);
void cra_ccf_tm_sched_free(void* sched);
// the contraction on tcgen05.mma, W streamed through tensor memory class by class (cra_ccf_um.cu; maxrin 256)
bool cra_ccf_um_supported(int log2n);
int cra_ccf_um_num_tiles(int R);
size_t cra_ccf_um_refimg_bytes(int max_refs, int nch);
int cra_ccf_um_pack_refs(const unsigned char* refspec, int R, const CraFragTab& frag, unsigned char* img, cudaStream_t st);
int cra_launch_ccf_um(const unsigned char* spec, int nrows, const unsigned char* refimg, int R, const CraRingTab& htab,
                      const CraFragTab& frag, const std::vector<int>& h_koff, const float2* twid, CraCand* cand,
                      int ntile_n, const float2* norm, const float* tref, cudaStream_t st);
int cra_launch_ccf(const float* spec, int nrows, const float* refspec, int R, const CraRingTab* tab,
                   const CraRingTab& htab, const float2* twid, CraCand* cand, int ntile_n, cudaStream_t st);
int cra_ccf_tile_n();
// twd: [maxrin] (cos, sin)(2 pi j / maxrin) in double; norm / tref: deferred Normalize_ring (fragment format), else null
int cra_launch_finalize(const float* spec, const float* refspec, int R, const CraRingTab* tab, const CraRingTab& htab,
                        const CraCand* cand, int ntile_n, CraRowMap map, CraResult* out, int fmt, const CraFragTab& frag,
                        const double2* twd, const float2* norm, const float* tref, cudaStream_t st);
int cra_launch_ccf_curves(const float* spec, int row, const float* refspec, int ref, const CraRingTab* tab,
                          const CraRingTab& htab, float* q, float* t, int fmt, const CraFragTab& frag, cudaStream_t st);
void cra_ccf_twiddles(int log2n, std::vector<float2>& tw);
int cra_launch_rotsum(const float* images, int nx, int p0, int n, const float4* params, const int* iref,
                      long global_offset, float* sums, float* counts, float* out_images, cudaStream_t st);
int cra_fp32_peak(double* tf_ffma, double* tf_ffma2);
// references from class sums, tangent low-pass of nx x nx images in place (cra_refavg.cu)
int cra_launch_class_average(const float* sums, const float* counts, float* refs, int R, int nx, cudaStream_t st);
// scratch: global buffer for boxes whose spectra exceed shared memory (cra_dft_scratch_elems float2 per image; else unused)
size_t cra_dft_scratch_elems(int nx, int nsh, int nspec);
int cra_launch_tanl_filter(float* imgs, int n, int nx, float fl, float aa, float2* scratch, cudaStream_t st);
// reference update on the device (cra_refupdate.cu): class averages + ring-binned FSC sums; filt_tanl + centring
int cra_launch_class_fsc(const float* sums, const float* counts, float* refs, const short* shell, const float* mask, int R,
                         int nx, int nsh, int masked, int min_members, int write_avg, float avg_div, double* fsc_out,
                         float2* scratch, cudaStream_t st);
int cra_launch_filter_center(float* imgs, int n, int nx, float fl, float aa, int mode, float sx, float sy, float* cs_out,
                             float2* scratch, cudaStream_t st);
