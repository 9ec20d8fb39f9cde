// cra_compat.cu -- layer 2 of include/cryo_ralib.h: the symbols the reference's Python
// drivers bind from cuda/gpu_aln_pack.so (cuda/gpu_aln_noref.h:52-113; definitions
// cuda/gpu_aln_noref.cu:119-782), re-implemented on top of the cra_* core so that the
// reference's driver logic (test_mref_gpu_align.py:365-449, test_reffree.py:270-430,
// test_mref_cheng_yu_bdb_cuda.py:546-556) carries over unchanged while the numbers
// follow EMAN2/Sphire multiref_polar_ali_2d / ormq semantics.
//
// Like the reference this layer is a process-global singleton (gpu_aln_noref.cu:35-51)
// and reports failure by printing and aborting the call with a null result instead of
// exit(): callers that want status codes use the cra_* layer.
#include "cra_common.cuh"
#include <math.h>
#include <stdio.h>
#include <string.h>

namespace {

struct Legacy {
    CraCtx* ctx = nullptr;
    AlignConfig cfg{};
    unsigned num_particles = 0;
    int device = 0;
    AlignParam* params = nullptr;     // managed, [num_particles]
    float xr = 0.f, step = 1.f;
    float* h_stage = nullptr; size_t stage_n = 0;       // pinned gather buffer for pre_align_fetch
    float* d_trans = nullptr; size_t trans_n = 0;       // transformed images of the last run
    float* m_sums = nullptr;                            // managed [2R][nx][nx]: even block, odd block
    int* m_counts = nullptr;                            // managed [R]
    std::vector<CraSearch> search;
    std::vector<CraResult> res;
    std::vector<float> par;
    std::vector<int> iref;
    std::vector<int> cid;             // class of every particle (ref_free_alignment_2D_init), empty otherwise
} g;

void fail(const char* where)
{
    fprintf(stderr, "[cryo_ralib] %s failed: %s\n", where, cra_last_error());
}

CraConfig core_config(const AlignConfig* a, unsigned capacity)
{
    CraConfig c{};
    c.nx = (int)a->img_dim;
    c.ir = 1; c.ou = (int)a->ring_num; c.rs = 1;
    c.max_particles = (int)capacity;
    c.max_refs = (int)a->ref_num;
    c.max_range = fmaxf(a->shift_rng_x, a->shift_rng_y);
    c.step = a->shift_step;
    c.normalize_ring = 1;
    c.row_batch = 0;
    return c;
}

// search_range after the driver's swap (sp_alignment.search_range; test_mref.py:195-198)
void search_range(int n, int radius, double shift, double range, float* l, float* r)
{
    const int cn = n / 2 + 1;
    double ql = cn + shift - radius - 2, qe = n - cn - shift - radius;
    if (ql < 0) ql = 0;
    if (qe < 0) qe = 0;
    *l = (float)(ql < range ? ql : range);
    *r = (float)(qe < range ? qe : range);
}

// One alignment pass over particles [start,stop) of the fetched batch.
//   multiref: Normalize_ring on, out-of-range shifts reset (test_mref.py:190-193)
//   !multiref: ormq semantics, shifts clamped (ali2d_single_iter)
int run_alignment(int start, int stop, bool multiref, bool bound = false)
{
    if (!g.ctx) { cra_set_error("pre_align_init has not been called"); return 1; }
    const int n = stop - start;
    if (n <= 0) return 0;
    if (start < 0 || (unsigned)stop > g.num_particles) { cra_set_error("index range outside the particle list"); return 1; }
    const int nx = (int)g.cfg.img_dim, ou = (int)g.cfg.ring_num;
    const int cnx = nx / 2 + 1;
    const int mashi = cnx - ou - 2;
    g.search.resize(n); g.res.resize(n);
    for (int i = 0; i < n; ++i) {
        AlignParam& p = g.params[start + i];
        double sxi = p.shift_x, syi = p.shift_y;
        if (multiref) {
            if (fabs(sxi) > mashi || fabs(syi) > mashi) { sxi = 0.0; syi = 0.0; }
        } else {
            sxi = fmin(fmax(sxi, -(double)mashi), (double)mashi);
            syi = fmin(fmax(syi, -(double)mashi), (double)mashi);
        }
        p.shift_x = (float)sxi; p.shift_y = (float)syi;
        CraSearch& s = g.search[i];
        s.cx = (float)(cnx + sxi); s.cy = (float)(cnx + syi);
        search_range(nx, ou, sxi, g.xr, &s.xl, &s.xr);
        search_range(nx, ou, syi, g.xr, &s.yl, &s.yr);
    }
    if (cra_set_normalize_ring(g.ctx, multiref ? 1 : 0)) return 1;
    if (cra_set_step(g.ctx, g.step)) return 1;
    if (bound) {
        if (g.cid.size() != g.num_particles) { cra_set_error("ref_free_alignment_2D_init has not been called"); return 1; }
        if (cra_align_bound(g.ctx, 0, n, g.search.data(), g.cid.data() + start, g.res.data())) return 1;
    } else if (cra_align(g.ctx, 0, n, g.search.data(), g.res.data())) return 1;
    for (int i = 0; i < n; ++i) {
        AlignParam& p = g.params[start + i];
        const CraResult& r = g.res[i];
        p.sbj_id = start + i;
        p.ref_id = r.iref;
        p.shift_x -= r.sx;           // r.sx = -ix: the centre moves by +ix (gpu_aln_noref.cu:1476)
        p.shift_y -= r.sy;
        p.angle = r.ang;
        p.mirror = r.mirror != 0;
    }
    return 0;
}

// EMAN2 parameters of particle i from its AlignParam: the a19 conversion
// (test_mref_gpu_align.py:578-588), identical to combine_params2(0,-sxi,-syi,0, ang,sxs,sys,m).
void eman_params(const AlignParam& p, float* out4)
{
    const double a = (double)p.angle * M_PI / 180.0;
    const double c = cos(a), s = -sin(a);
    const double sxn = -(double)p.shift_x, syn = -(double)p.shift_y;
    out4[0] = p.angle;
    out4[1] = (float)(sxn * c - syn * s);
    out4[2] = (float)(sxn * s + syn * c);
    out4[3] = p.mirror ? 1.0f : 0.0f;
}

}  // namespace

extern "C" void print_gpu_info(const unsigned int device_idx)
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, (int)device_idx) != cudaSuccess) { printf("no CUDA device %u\n", device_idx); return; }
    size_t fr = 0, tot = 0;
    cudaSetDevice((int)device_idx);
    cudaMemGetInfo(&fr, &tot);
    printf("GPU[%u]: %s, sm_%d%d, %d SMs, %.1f GB total / %.1f GB free, L2 %.0f MB, smem/SM %zu KB\n",
           device_idx, p.name, p.major, p.minor, p.multiProcessorCount, tot / 1073741824.0, fr / 1073741824.0,
           p.l2CacheSize / 1048576.0, p.sharedMemPerMultiprocessor / 1024);
}

// Memory model of this engine (replaces the CCF-table estimate of gpu_aln_noref.cu:234-349):
// resident images + one row batch of spectra + candidates + refs + sums.
extern "C" bool pre_align_size_check(const unsigned int num_particles, const AlignConfig* cfg,
                                     const unsigned int cuda_device_id, const float request, const bool verbose)
{
    (void)num_particles;
    if (!cfg) return false;
    if (cudaSetDevice((int)cuda_device_id) != cudaSuccess) return false;
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return false;
    const double npix = (double)cfg->img_dim * cfg->img_dim;
    const double lcirc_est = 2.0 * 2.0 * M_PI * cfg->ring_num * (cfg->ring_num + 1) / 2.0 * 1.5;   // <= 2x oversampled rings
    const double k = cfg->shift_step > 0 ? floor(fmaxf(cfg->shift_rng_x, cfg->shift_rng_y) / cfg->shift_step) : 0;
    const double smax = (2 * k + 1) * (2 * k + 1);
    double rows = fmin((double)cfg->sbj_num * smax, fmax(smax, 2147483648.0 / (lcirc_est * 4)));
    double need = cfg->sbj_num * npix * 4 * 2      /* images + transformed images */
                + rows * lcirc_est * 4 + rows * ((cfg->ref_num + 3) / 4) * 8
                + cfg->ref_num * (npix * 4 * 5 + lcirc_est * 4) + (64 << 20);
    const bool ok = need <= (double)fr * request;
    if (verbose) printf("pre_align_size_check: need %.2f GB, free %.2f GB x %.2f -> %s\n", need / 1073741824.0,
                        fr / 1073741824.0, request, ok ? "fits" : "does not fit");
    return ok;
}

extern "C" void gpu_clear(void)
{
    if (g.ctx) { cra_destroy(g.ctx); g.ctx = nullptr; }
    if (g.params) { cudaFree(g.params); g.params = nullptr; }
    if (g.h_stage) { cudaFreeHost(g.h_stage); g.h_stage = nullptr; g.stage_n = 0; }
    if (g.d_trans) { cudaFree(g.d_trans); g.d_trans = nullptr; g.trans_n = 0; }
    if (g.m_sums) { cudaFree(g.m_sums); g.m_sums = nullptr; }
    if (g.m_counts) { cudaFree(g.m_counts); g.m_counts = nullptr; }
    g.cid.clear();
    g.num_particles = 0;
}

extern "C" AlignParam* pre_align_init(const unsigned int num_particles, const AlignConfig* cfg,
                                      const unsigned int cuda_device_id)
{
    if (!cfg || num_particles == 0) { fprintf(stderr, "[cryo_ralib] pre_align_init: bad arguments\n"); return nullptr; }
    gpu_clear();
    g.cfg = *cfg; g.num_particles = num_particles; g.device = (int)cuda_device_id;
    g.xr = fmaxf(cfg->shift_rng_x, cfg->shift_rng_y); g.step = cfg->shift_step;
    unsigned cap = cfg->sbj_num ? cfg->sbj_num : num_particles;
    if (cap > num_particles) cap = num_particles;
    CraConfig cc = core_config(cfg, cap);
    if (cra_create(&cc, g.device, &g.ctx)) { fail("pre_align_init"); g.ctx = nullptr; return nullptr; }
    if (cudaMallocManaged(&g.params, sizeof(AlignParam) * num_particles) != cudaSuccess) { fprintf(stderr, "[cryo_ralib] managed alloc failed\n"); gpu_clear(); return nullptr; }
    const size_t npix = (size_t)cfg->img_dim * cfg->img_dim;
    if (cudaMallocManaged(&g.m_sums, sizeof(float) * 2 * cfg->ref_num * npix) != cudaSuccess ||
        cudaMallocManaged(&g.m_counts, sizeof(int) * cfg->ref_num) != cudaSuccess) { fprintf(stderr, "[cryo_ralib] managed alloc failed\n"); gpu_clear(); return nullptr; }
    for (unsigned i = 0; i < num_particles; ++i) {
        g.params[i].sbj_id = (int)i; g.params[i].ref_id = -1; g.params[i].shift_x = 0; g.params[i].shift_y = 0;
        g.params[i].angle = 0; g.params[i].mirror = false;
    }
    return g.params;
}

extern "C" void pre_align_fetch(const float** img_data, const unsigned int img_num, const char* batch_type)
{
    if (!g.ctx || !img_data || !batch_type) { fprintf(stderr, "[cryo_ralib] pre_align_fetch: not initialised\n"); return; }
    const size_t npix = (size_t)g.cfg.img_dim * g.cfg.img_dim;
    const size_t need = (size_t)img_num * npix;
    if (need > g.stage_n) {
        if (g.h_stage) cudaFreeHost(g.h_stage);
        if (cudaMallocHost(&g.h_stage, need * sizeof(float)) != cudaSuccess) { fprintf(stderr, "[cryo_ralib] pinned alloc failed\n"); g.h_stage = nullptr; g.stage_n = 0; return; }
        g.stage_n = need;
    }
    // gather of the borrowed per-image host pointers into the pinned staging buffer, all host threads
#pragma omp parallel for schedule(static) num_threads(cra_host_threads())
    for (long i = 0; i < (long)img_num; ++i) memcpy(g.h_stage + (size_t)i * npix, img_data[i], npix * sizeof(float));
    int rc;
    if (strcmp(batch_type, "sbj_batch") == 0) rc = cra_upload_particles(g.ctx, g.h_stage, 0, (int)img_num, 0);
    else if (strcmp(batch_type, "ref_batch") == 0) rc = cra_set_refs(g.ctx, g.h_stage, (int)img_num, 0);
    else { fprintf(stderr, "[cryo_ralib] pre_align_fetch: unknown batch type '%s'\n", batch_type); return; }
    if (rc) fail("pre_align_fetch");
}

extern "C" void reset_shifts(const float shift_range, const float shift_step)
{
    if (!g.ctx) { fprintf(stderr, "[cryo_ralib] reset_shifts: not initialised\n"); return; }
    const float cap = fmaxf(g.cfg.shift_rng_x, g.cfg.shift_rng_y);
    if (shift_step <= 0.f || floorf(shift_range / shift_step) > floorf(cap / g.cfg.shift_step)) {
        fprintf(stderr, "[cryo_ralib] reset_shifts: grid larger than the one configured at init\n"); return;
    }
    g.xr = shift_range; g.step = shift_step;
}

static int transform_batch(int start, int stop, bool want_images, bool want_sums)
{
    const int n = stop - start;
    const size_t npix = (size_t)g.cfg.img_dim * g.cfg.img_dim;
    g.par.resize((size_t)4 * n); g.iref.resize(n);
    for (int i = 0; i < n; ++i) { eman_params(g.params[start + i], &g.par[4 * i]); g.iref[i] = g.params[start + i].ref_id; }
    if (want_images) {
        if ((size_t)n * npix > g.trans_n) {
            if (g.d_trans) cudaFree(g.d_trans);
            if (cudaMalloc(&g.d_trans, (size_t)n * npix * sizeof(float)) != cudaSuccess) { cra_set_error("transformed-image alloc failed"); g.d_trans = nullptr; g.trans_n = 0; return 1; }
            g.trans_n = (size_t)n * npix;
        }
        if (cra_transform_dev(g.ctx, 0, n, g.par.data(), g.d_trans)) return 1;      // stays on the device
    }
    if (want_sums) {
        const int R = (int)g.cfg.ref_num;
        if (cra_zero_sums(g.ctx)) return 1;
        if (cra_accumulate(g.ctx, 0, n, g.par.data(), g.iref.data(), start)) return 1;
        std::vector<float> s((size_t)R * 2 * npix), cnt(R);
        if (cra_get_sums(g.ctx, s.data(), cnt.data())) return 1;
        // reference layout: [2R][nx][nx] = all even sums, then all odd sums (gpu_aln_noref.cu:1232-1274)
        for (int r = 0; r < R; ++r) {
            memcpy(g.m_sums + (size_t)r * npix, &s[((size_t)r * 2 + 0) * npix], npix * sizeof(float));
            memcpy(g.m_sums + (size_t)(R + r) * npix, &s[((size_t)r * 2 + 1) * npix], npix * sizeof(float));
            g.m_counts[r] = (int)(cnt[r] + 0.5f);
        }
    }
    return 0;
}

extern "C" void* mref_align_run(const int start_idx, const int stop_idx)
{
    if (run_alignment(start_idx, stop_idx, true) || transform_batch(start_idx, stop_idx, true, false)) { fail("mref_align_run"); return nullptr; }
    return g.d_trans;
}

extern "C" float* mref_align_run_m(const int start_idx, const int stop_idx)
{
    if (run_alignment(start_idx, stop_idx, true) || transform_batch(start_idx, stop_idx, false, true)) { fail("mref_align_run_m"); return nullptr; }
    return g.m_sums;
}

extern "C" int* get_num_ref(void) { return g.m_counts; }

extern "C" void pre_align_run(const int start_idx, const int stop_idx)
{
    if (run_alignment(start_idx, stop_idx, false)) fail("pre_align_run");
}

extern "C" void* pre_align_run_m(const int start_idx, const int stop_idx)
{
    if (run_alignment(start_idx, stop_idx, false) || transform_batch(start_idx, stop_idx, true, false)) { fail("pre_align_run_m"); return nullptr; }
    return g.d_trans;
}

// ---- gpu_isac's class-bound reference-free alignment (gpu_aln_noref.cu:559-782) ------------------

extern "C" bool ref_free_alignment_2D_size_check(const AlignConfig* cfg, const unsigned int cuda_device_id,
                                                 const float request, const bool verbose)
{
    return cfg ? pre_align_size_check(cfg->sbj_num, cfg, cuda_device_id, request, verbose) : false;
}

extern "C" AlignParam* ref_free_alignment_2D_init(const AlignConfig* cfg, const float** sbj_data_list,
                                                  const float** ref_data_list, const int* sbj_cid_list,
                                                  const unsigned int cuda_device_id)
{
    if (!cfg || !sbj_data_list || !ref_data_list || !sbj_cid_list) { fprintf(stderr, "[cryo_ralib] ref_free_alignment_2D_init: bad arguments\n"); return nullptr; }
    for (unsigned i = 0; i < cfg->sbj_num; ++i)
        if (sbj_cid_list[i] < 0 || (unsigned)sbj_cid_list[i] >= cfg->ref_num) { fprintf(stderr, "[cryo_ralib] ref_free_alignment_2D_init: class id outside the reference list\n"); return nullptr; }
    if (!pre_align_init(cfg->sbj_num, cfg, cuda_device_id)) return nullptr;
    pre_align_fetch(sbj_data_list, cfg->sbj_num, "sbj_batch");
    pre_align_fetch(ref_data_list, cfg->ref_num, "ref_batch");
    g.cid.assign(sbj_cid_list, sbj_cid_list + cfg->sbj_num);
    for (unsigned i = 0; i < cfg->sbj_num; ++i) g.params[i].ref_id = sbj_cid_list[i];      // gpu_aln_noref.cu:598-599
    return g.params;
}

extern "C" void ref_free_alignment_2D(void)
{
    const int n = (int)g.num_particles;
    if (run_alignment(0, n, false, true)) { fail("ref_free_alignment_2D"); return; }
    // apply_alignment_param + fetch_averages (gpu_aln_noref.cu:771-775): class averages of the transformed images
    g.par.resize((size_t)4 * n);
    for (int i = 0; i < n; ++i) eman_params(g.params[i], &g.par[4 * i]);
    if (cra_zero_sums(g.ctx) || cra_accumulate(g.ctx, 0, n, g.par.data(), g.cid.data(), 0) || cra_refs_from_sums(g.ctx, 0))
        fail("ref_free_alignment_2D");
}

extern "C" void ref_free_alignment_2D_filter_references(const float cutoff_freq, const float falloff)
{
    if (!g.ctx) { fprintf(stderr, "[cryo_ralib] ref_free_alignment_2D_filter_references: not initialised\n"); return; }
    if (cra_filter_refs(g.ctx, cutoff_freq, falloff, 0)) fail("ref_free_alignment_2D_filter_references");
}
