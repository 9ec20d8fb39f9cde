/*
 * cryo_ralib.h -- C ABI of libcryo_ralib.so (B200 / sm_100a engine for cryo-EM
 * 2D multi-reference and reference-free alignment).
 *
 * Two layers, both plain `extern "C"` (pointers + sizes, no torch / C++ types):
 *
 *  (1) cra_*  : status-returning, context-based core.  Every call returns 0 on
 *               success; cra_last_error() gives the message.  Nothing exits
 *               the process, nothing falls back to the CPU.
 *  (2) legacy : the exact symbols the reference's Python drivers bind from
 *               cuda/gpu_aln_pack.so (cuda/gpu_aln_noref.h:52-113, struct
 *               layouts cuda/gpu_aln_common.h:62-83, ctypes mirrors
 *               test_mref_gpu_align.py:112-131), implemented on top of (1)
 *               with EMAN2/Sphire multiref_polar_ali_2d semantics.
 *
 * Conventions: images are row-major float32 [n][nx][nx] (square);  pixel
 * coordinates follow EMAN2/SPIDER: centre cnx = nx/2+1 (1-based).
 */
#ifndef CRYO_RALIB_H
#define CRYO_RALIB_H

#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ (1) core */

typedef struct CraCtx CraCtx;

/* Replaces the reference's AlignConfig (gpu_aln_common.h:62-75) and adds the
 * EMAN2 switches it lacks: ir/rs, separate yr, float step. */
typedef struct CraConfig {
    int   nx;             /* image edge (square)                                     */
    int   ir, ou, rs;     /* first ring, last ring, ring step   (Numrinit, mode "F") */
    int   max_particles;  /* capacity of the resident particle stack                 */
    int   max_refs;       /* capacity of the reference stack                         */
    float max_range;      /* largest |xr|,|yr| that will be requested                */
    float step;           /* translational step ts (may be < 1)                      */
    int   normalize_ring; /* 1: multiref_polar_ali_2d (Normalize_ring); 0: ormq      */
    int   row_batch;      /* particle-shift rows per device batch; 0 = auto          */
} CraConfig;

/* Per-particle search request: polar centre (1-based, = cnx+sxi, cny+syi) and
 * the permissible window [left,right] per axis as search_range returns it after
 * the driver's swap (test_mref.py:195-201).  Grid positions are
 * ix = j*step, j = -int(xl/step) .. int(xr/step); same for y.           */
typedef struct CraSearch {
    float cx, cy;
    float xl, xr, yl, yr;
} CraSearch;

/* Result of Util.multiref_polar_ali_2d for one particle (test_mref.py:200):
 * [ang, sxs, sys, mirror, iref, peak]; sx,sy are the raw grid offsets (-ix,-iy). */
typedef struct CraResult {
    float ang, sxs, sys;
    int   mirror;
    int   iref;
    float peak;
    float sx, sy;
} CraResult;

int  cra_create(const CraConfig* cfg, int device, CraCtx** out);
int  cra_destroy(CraCtx* ctx);
const char* cra_last_error(void);

/* Ring geometry the context was built with (host copies; numr = nring triplets
 * radius, 1-based offset, length  -- Numrinit, test_mref.py:145). */
int  cra_ring_info(CraCtx* ctx, int* nring, int* lcirc, int* maxrin, int* numr_out /*3*nring or NULL*/);

/* Upload n particles (host, [n][nx][nx]) into slots [first, first+n).  If
 * subtract_mask_mean != 0 the mean under model_circle(ou) is subtracted on the
 * device: normalize.mask no_sigma=0 (test_mref.py:188).  Replaces
 * pre_align_fetch(...,"sbj_batch") (gpu_aln_noref.cu:362-380).              */
int  cra_upload_particles(CraCtx* ctx, const float* host_images, int first, int n, int subtract_mask_mean);
/* Same, source already on this device. */
int  cra_upload_particles_dev(CraCtx* ctx, const float* dev_images, int first, int n, int subtract_mask_mean);

/* Asynchronous variant: the copy is queued on the context's copy stream and returns at once;
 * cra_align / cra_accumulate / cra_transform on a particle range wait, on the device, only for the
 * uploads that overlap it, so the upload of the next chunk runs under the alignment of the current
 * one.  The mask-mean subtraction of an uploaded range runs on the main stream in front of its first
 * consumer (or in cra_upload_wait), once for the whole range: a borrowed cra_device_images_ptr shows
 * the raw pixels until then.  host_images must be pinned and stay valid until cra_upload_wait (or a
 * later call that consumed the range) returns.                                                   */
int  cra_upload_particles_async(CraCtx* ctx, const float* host_images, int first, int n, int subtract_mask_mean);
int  cra_upload_wait(CraCtx* ctx);

/* Upload R references and prepare them: optional normalize.mask(no_sigma=1),
 * Polar2Dm at (cnx,cny), Frngs, Applyws (test_mref.py:170-175).  Replaces
 * pre_align_fetch(...,"ref_batch").                                         */
int  cra_set_refs(CraCtx* ctx, const float* host_refs, int R, int normalize_mask);

/* Multi-reference alignment of particles [start,stop): the a7 hot path
 * (Polar2Dm -> Normalize_ring -> Frngs -> Crosrng_ms -> best).  search and out
 * are host arrays indexed from 0 for particle `start`.                      */
int  cra_align(CraCtx* ctx, int start, int stop, const CraSearch* search, CraResult* out);

/* Class-bound alignment: particle p is matched against reference class_of[p] only, over the same
 * shift window and with the same arithmetic as cra_align.  This is the alignment step of gpu_isac's
 * ref_free_alignment_2D (cuda/gpu_aln_noref.cu:743-770: every image of a class against that class's
 * average); class_of is the reference's sbj_cid_list (gpu_aln_noref.cu:566-571), host int[stop-start].
 * out[p].iref = class_of[p].                                                                     */
int  cra_align_bound(CraCtx* ctx, int start, int stop, const CraSearch* search, const int* class_of, CraResult* out);
/* References rebuilt ON THE DEVICE from the accumulated class sums: ref[r] = (even[r] + odd[r]) / count[r]
 * (classes without members keep their reference), then prepared like cra_set_refs.  Replaces
 * BatchHandler::fetch_averages (gpu_aln_noref.cu:772-775).  After N > 1 ranks allreduce the sums buffer
 * every rank gets the same references.                                                          */
int  cra_refs_from_sums(CraCtx* ctx, int normalize_mask);
/* Tangent low-pass of the current references in place (EMAN2 filt_tanl semantics; replaces
 * ref_free_alignment_2D_filter_references, gpu_aln_noref.cu:777-814), then prepared again.     */
int  cra_filter_refs(CraCtx* ctx, float cutoff_freq, float falloff, int normalize_mask);
/* Current reference images [R][nx][nx] to the host (as prepared: after normalize.mask when that was asked). */
int  cra_get_refs(CraCtx* ctx, float* host_refs);

/* The per-iteration reference update ON THE DEVICE (the rank-0 section of the reference's loop, test_mref.py:238-286;
 * reference-free twin test_reffree.py:695-755).  The class sums never leave the GPU; the host keeps what the reference
 * keeps in Python: the class average of the FSC curves and the 2-parameter tangent fit (cra_fit_tanh).
 *
 * cra_class_fsc: for every class r with counts[r] >= min_members (test_mref.py:244: 4):
 *   write_avg != 0: reference slot r <- (even[r] + odd[r]) / (avg_div > 0 ? avg_div : counts[r])   (test_mref.py:255-256;
 *                   the reference-free driver divides by the particle count, test_reffree.py:697)
 *   fsc_out[r][i] = fsc(even[r], odd[r]) at shell i (EMData::calc_fourier_shell_correlation, w = 1; test_mref.py:254);
 *                   masked != 0: fsc_mask -- in-mask mean removed and masked first (test_reffree.py:708).
 *   Classes below min_members get a zero curve and keep their slot: the caller reseeds them with cra_put_ref.
 *   nshell = nx/2 + 1; freq_out [nshell], n_out [nshell] are the other two columns sp_statistics.fsc returns;
 *   counts_out [R].  fsc_out == NULL only reports nshell.
 * cra_put_ref: overwrite one reference slot from the host (the <4-member reseed, test_mref.py:244-249).
 * cra_filter_center_refs: every reference <- filt_tanl(ref, cutoff, falloff), then mode 1: center_2D(., 1) = phase_cog +
 *   fshift(-cs) (sp_user_functions.ref_ali2d, test_mref.py:273-276); mode 2: fshift(ref, -sx, -sy) (test_reffree.py:741-745);
 *   mode 0: no shift; then normalize.mask(no_sigma = 1) when normalize_mask (test_mref.py:284).  cs_out [R][2] (may be NULL).
 * cra_prepare_refs: the reference preparation of the next iteration (test_mref.py:170-175) on the references already
 *   on the device -- cra_set_refs without the upload.                                                             */
int  cra_class_fsc(CraCtx* ctx, int masked, int min_members, int write_avg, float avg_div, int* nshell,
                   double* freq_out, double* fsc_out, double* n_out, float* counts_out);
int  cra_put_ref(CraCtx* ctx, int iref, const float* host_img);
int  cra_filter_center_refs(CraCtx* ctx, float cutoff_freq, float falloff, int mode, float sx, float sy,
                            int normalize_mask, float* cs_out);
int  cra_prepare_refs(CraCtx* ctx, int normalize_mask);

/* Host bookkeeping of the reference's per-particle Python loop, batched (no device work):
 * cra_mref_search_request = get_params2D -> inverse_transform2 -> mashi reset -> search_range x2
 * (test_mref.py:184-198); params [n][4] double (alpha, sx, sy, mirror) is reset in place where the
 * reference resets it.  cra_compose_result = combine_params2(0,-sxi,-syi,0, ang,sxs,sys,mirror)
 * (test_mref.py:206) -> params_out [n][4] double.                                              */
int  cra_mref_search_request(int n, double* params, int nx, int ou, double xr, double yr,
                             CraSearch* search, double* sxi_out, double* syi_out);
int  cra_compose_result(int n, const double* sxi, const double* syi, const CraResult* res, double* params_out);
/* The reference-free twin (ali2d_single_iter, test_reffree.py:780-783): combine_params2(params, 0, -cs_x, -cs_y, 0)
 * -> inverse_transform2 -> shift clamped to +-mashi -> search_range x2.  params [n][4] is not modified.        */
int  cra_reffree_search_request(int n, const double* params, double cs_x, double cs_y, int nx, int ou,
                                double xr, double yr, CraSearch* search, double* sxi_out, double* syi_out);

/* sp_filter.fit_tanh's optimisation (host arithmetic, no device work): Nelder-Mead fit (sp_utilities.amoeba, same
 * simplex, tolerances 1e-4, 500 iterations at most) of 0.5 (tanh(c (f + fl)) - tanh(c (f - fl))), c = pi / (2 aa fl), to
 * target[i] = 2 FSC / (1 + FSC) at freq[i]; start (fl0, aa0), simplex scale (scale_fl, scale_aa).  The user
 * function ref_ali2d runs it for every reference update (test_mref.py:273-276, test_reffree.py:733).
 * out4 = fl, aa, -sum of squares, iterations.                                                            */
int  cra_fit_tanh(int n, const double* freq, const double* target, double fl0, double aa0,
                  double scale_fl, double scale_aa, double* out4);

/* rot_shift2D(img, alpha, sx, sy, mirror) + add into class sums
 * sums[iref][global_index % 2] and counts[iref] (test_mref.py:210-215).
 * params: host [n][4] float (alpha, sx, sy, mirror); iref: host int[n]; iref<0
 * skips the particle.  global_offset is the global index of particle `start`. */
int  cra_accumulate(CraCtx* ctx, int start, int stop, const float* params, const int* iref,
                    long global_offset);
/* The same with the parameters in double, as the reference's Python holds them (test_mref.py:206-210).  They are
 * rounded to float exactly where EMAN2 does it -- rot_scale_trans2D_background takes float arguments -- and the angle is
 * then converted to radians in double and rounded once (float ang = angDeg * M_PI / 180.0f), which pins the quadri
 * cell of every output pixel.                                                                                      */
int  cra_accumulate_d(CraCtx* ctx, int start, int stop, const double* params, const int* iref,
                      long global_offset);
int  cra_zero_sums(CraCtx* ctx);
/* Device pointers (owned by ctx) of the packed [R][2][nx][nx] f32 sums followed
 * by [R] f32 counts: one contiguous buffer so a single NCCL allreduce covers
 * both (replaces reduce_EMData_to_root x2R + mpi_reduce, test_mref.py:219-223). */
int  cra_sums_device_ptr(CraCtx* ctx, void** dev_ptr, size_t* n_floats);
int  cra_get_sums(CraCtx* ctx, float* host_sums /*[R][2][nx][nx]*/, float* host_counts /*[R]*/);

/* rot_shift2D only: transformed images of [start,stop) to a host buffer
 * (ref-free sum_oe / apply-transform export).                               */
int  cra_transform(CraCtx* ctx, int start, int stop, const float* params, float* host_out);
int  cra_transform_d(CraCtx* ctx, int start, int stop, const double* params, float* host_out);
/* The same into a caller-owned DEVICE buffer [stop-start][nx][nx] (what mref_align_run hands back,
 * gpu_aln_noref.cu:389-416: the transformed images never leave the GPU).                       */
int  cra_transform_dev(CraCtx* ctx, int start, int stop, const float* params, float* dev_out);

/* Stage-level entry points used by the parity tests. */
int  cra_polar_spectrum(CraCtx* ctx, int particle, float cx, float cy, float* host_out /*lcirc*/);
int  cra_ref_spectrum(CraCtx* ctx, int iref, float* host_out /*lcirc*/);
/* Spectrum (Normalize_ring applied) of row `row` of the LAST row batch cra_align processed, as the
 * production row kernel left it: rows run particle by particle in the visit order of
 * multiref_polar_ali_2d (y outer, x inner).  which_kernel: 1 = grouped row kernel, 0 = general. */
int  cra_batch_row_spectrum(CraCtx* ctx, int row, float* host_out /*lcirc*/, int* which_kernel);
int  cra_ccf_curves(CraCtx* ctx, int particle, float cx, float cy, int iref,
                    float* q_out /*maxrin*/, float* t_out /*maxrin*/);

/* Timing of the last cra_align call, measured with CUDA events on the engine's
 * stream: ms spent in the polar/FFT kernel, the CCF/peak kernel, the finalize
 * kernel, and the number of kernel launches and alignments evaluated.        */
typedef struct CraAlignStats {
    float ms_polar, ms_ccf, ms_final, ms_total;
    long  launches;
    long  alignments;   /* particle x reference x shift triples actually evaluated */
    long  rows;         /* particle x shift rows                                  */
} CraAlignStats;
int  cra_last_align_stats(CraCtx* ctx, CraAlignStats* out);
int  cra_set_timing(CraCtx* ctx, int enabled);
/* Switch between multiref_polar_ali_2d (1) and ormq (0) ring normalisation, and
 * change the translational step between calls (multi-step xr/ts schedules).      */
int  cra_set_normalize_ring(CraCtx* ctx, int enabled);
int  cra_set_step(CraCtx* ctx, float step);
int  cra_row_batch(CraCtx* ctx);
/* Borrowed device pointer of the resident particle stack [max_particles][nx][nx]
 * and the engine's CUDA stream (cudaStream_t), for zero-copy interop.            */
int  cra_device_images_ptr(CraCtx* ctx, void** dev_ptr);
void* cra_stream(CraCtx* ctx);
/* FP32 FMA throughput of this device measured with a dependent-chain-free
 * FFMA2 micro-kernel (TFLOP/s); the roofline denominator for the CCF kernel. */
int  cra_measure_fp32_peak(int device, double* tflops_ffma, double* tflops_ffma2);

/* --------------------------------------------------------------- (2) legacy */

/* gpu_aln_common.h:62-75 / test_mref_gpu_align.py:112-123 */
typedef struct AlignConfig {
    unsigned int sbj_num;
    unsigned int ref_num;
    unsigned int img_dim;
    unsigned int ring_num;   /* the drivers pass numr[-3] = ou here (test_mref_gpu_align.py:368) */
    unsigned int ring_len;   /* ignored: ring lengths follow Numrinit                            */
    float shift_step;
    float shift_rng_x;
    float shift_rng_y;
} AlignConfig;

/* gpu_aln_common.h:76-83 / test_mref_gpu_align.py:125-131 */
typedef struct AlignParam {
    int   sbj_id;
    int   ref_id;
    float shift_x;   /* accumulated polar-centre offset (= sxi+ix), as gpu_aln_noref.cu:1476 */
    float shift_y;
    float angle;     /* EMAN2 convention already; feed to the a19 conversion as is             */
    bool  mirror;
} AlignParam;

void        print_gpu_info(const unsigned int device_idx);                          /* gpu_aln_common.cu:165 */
bool        pre_align_size_check(const unsigned int num_particles, const AlignConfig* cfg,
                                 const unsigned int cuda_device_id, const float request,
                                 const bool verbose);                               /* gpu_aln_noref.cu:234 */
AlignParam* pre_align_init(const unsigned int num_particles, const AlignConfig* cfg,
                           const unsigned int cuda_device_id);                      /* gpu_aln_noref.cu:188 */
void        pre_align_fetch(const float** img_data, const unsigned int img_num,
                            const char* batch_type);                                /* gpu_aln_noref.cu:362 */
void        reset_shifts(const float shift_range, const float shift_step);          /* gpu_aln_noref.cu:119 */
void*       mref_align_run(const int start_idx, const int stop_idx);                /* gpu_aln_noref.cu:389 */
float*      mref_align_run_m(const int start_idx, const int stop_idx);              /* gpu_aln_noref.cu:419 */
int*        get_num_ref(void);                                                      /* gpu_aln_noref.cu:  get_num_ref */
void        pre_align_run(const int start_idx, const int stop_idx);                 /* gpu_aln_noref.cu:520 */
void*       pre_align_run_m(const int start_idx, const int stop_idx);               /* gpu_aln_noref.cu:489 */
void        gpu_clear(void);                                                        /* gpu_aln_noref.cu:141 */
/* gpu_isac's class-bound reference-free alignment (gpu_aln_noref.h:94-109; restype-only in the shipped
 * drivers).  sbj_cid_list[i] = class of particle i (runs of equal classes, as the reference
 * requires); every call of ref_free_alignment_2D() aligns each particle to its class average with ormq
 * semantics, accumulates the shifts in the returned AlignParam[] and rebuilds the averages on the device. */
AlignParam* ref_free_alignment_2D_init(const AlignConfig* aln_cfg, const float** sbj_data_list,
                                       const float** ref_data_list, const int* sbj_cid_list,
                                       const unsigned int cuda_device_id);          /* gpu_aln_noref.cu:559 */
bool        ref_free_alignment_2D_size_check(const AlignConfig* cfg, const unsigned int cuda_device_id,
                                             const float request, const bool verbose); /* gpu_aln_noref.cu:625 */
void        ref_free_alignment_2D(void);                                            /* gpu_aln_noref.cu:743 */
void        ref_free_alignment_2D_filter_references(const float cutoff_freq, const float falloff); /* gpu_aln_noref.cu:777 */

#ifdef __cplusplus
}
#endif
#endif /* CRYO_RALIB_H */
