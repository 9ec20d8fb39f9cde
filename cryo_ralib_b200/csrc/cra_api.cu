// cra_api.cu -- the status-returning C-ABI core (include/cryo_ralib.h, layer 1).
// Owns the device buffers (the reference's BatchHandler + CcfResultTable,
// cuda/gpu_aln_noref.cu:1500-2399, minus the CCF table), builds the ring / sampling
// tables (Sphire Numrinit, ringwe; EMAN2 alrl_ms geometry; test_mref.py:145-146) and
// sequences the kernels on one CUDA stream.  No exit(), no CPU fallback.
#include "cra_common.cuh"
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <mutex>
#include <map>
#include <utility>

static thread_local std::string g_err;
void cra_set_error(const std::string& msg) { g_err = msg; }
extern "C" const char* cra_last_error(void) { return g_err.c_str(); }

int cra_ensure_dyn_smem(const void* func, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> done;
    int dev = 0;
    CRA_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = done[std::make_pair(dev, func)];
    if (bytes > cur) {
        CRA_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cur = bytes;
    }
    return 0;
}

struct CraCtx {
    CraConfig cfg{};
    int device = 0;
    cudaStream_t st = nullptr;
    cudaStream_t st_copy = nullptr;                            // asynchronous particle uploads
    struct Pending { int first, n, sub; cudaEvent_t ev; };       // sub: 1 mask mean still to be subtracted, 0 to be measured (on the main stream), -1 superseded
    std::vector<Pending> pending;                              // uploads not yet ordered before the main stream
    std::vector<cudaEvent_t> ev_pool;
    int nx = 0, npix = 0, R = 0;
    CraRingTab htab{};
    std::vector<int> numr;
    int smax = 0;            // most rows one particle can have
    int row_batch = 0;
    int ntile_n_max = 0;
    bool timing = false;
    CraAlignStats stats{};
    // device
    CraRingTab* d_tab = nullptr;
    float4* d_samp = nullptr; float* d_sampw = nullptr;
    float2* d_twf = nullptr;  float2* d_twi = nullptr;
    double2* d_twd = nullptr;    // (cos, sin)(2 pi j / maxrin) in double (finalize_kernel)
    int* d_items = nullptr; CraPolarItems items{};
    float* d_mask = nullptr;
    float* d_dc = nullptr;       // [max_particles] in-mask mean a particle uploaded WITHOUT mean subtraction still carries
    float* d_images = nullptr; float* d_refs = nullptr; float* d_refspec = nullptr;
    float* d_spec = nullptr; CraCand* d_cand = nullptr;
    float* d_sums = nullptr;     // [R][2][npix] + [R]
    // per-call staging (grown on demand)
    size_t cap_meta = 0;         // particles
    char* h_meta = nullptr; char* d_meta = nullptr; size_t meta_bytes = 0;
    CraResult* d_res = nullptr; CraResult* h_res = nullptr;
    float4* d_par = nullptr; int* d_iref = nullptr; float4* h_par = nullptr; int* h_iref = nullptr; size_t cap_par = 0;
    float* d_tmpimg = nullptr; size_t cap_tmpimg = 0;
    float* d_curves = nullptr;
    float2* h_group = nullptr;   // pinned staging of one 4-row spectrum group (test entry points)
    std::vector<cudaEvent_t> ev;
    // contraction path: CRA_FMT_FRAG = tensor-core kernel on split-bf16 fragments (default),
    // CRA_FMT_F32 = FP32 FMA kernel on the float2 spectrum (CRA_CCF=simt)
    int fmt = CRA_FMT_FRAG;
    bool use_tm = true;          // W staged in tensor memory (cra_ccf_tm.cu); CRA_CCF=mma keeps it in shared memory
    void* tm_sched = nullptr;    // work lists of the tensor-memory CCF kernel (cra_ccf_tm.cu)
    bool use_um = false;         // contraction on tcgen05.mma (cra_ccf_um.cu); CRA_CCF=um, maxrin 256
    unsigned char* d_refimg = nullptr;   // UMMA reference operand images
    CraFragTab frag{};
    std::vector<int> h_koff, h_chunk_k;
    int* d_fragtab = nullptr;
    size_t row_bytes = 0;        // device spectrum bytes of one row in the active format
    CraGroupPlan plan{};         // grouped row kernel (cra_polar_grp.cu); plan.rmax == 0: unavailable
    void* d_plan = nullptr;
    bool use_group = true;       // CRA_POLAR=general forces the general kernel
    float2* d_dft = nullptr; size_t dft_elems = 0;             // reference-update scratch (large boxes only)
    int rpb = 2, gimg_rows = 0, gimg_one = 0;   // general row kernel: rows per CTA; image taps from global memory (large boxes)
    int last_rows = 0, last_group = 0;   // rows / kernel of the last batch (cra_batch_row_spectrum)
    // second pipeline lane (align_impl): odd row batches run on their own stream with their own spectrum / candidate
    // buffers, so that the tail of one batch's kernels is filled by the other batch's CTAs
    cudaStream_t st2 = nullptr;
    float* d_spec2 = nullptr; CraCand* d_cand2 = nullptr; float2* d_norm2 = nullptr;
    cudaEvent_t ev_lane[2] = {nullptr, nullptr};
    const float* last_spec = nullptr; const float2* last_norm = nullptr;     // buffers of the last batch
    float2* d_norm = nullptr;    // [row_batch] deferred Normalize_ring (avg, 1/sigma), cra_common.cuh
    float* d_tref = nullptr;     // [max_refs]  sum_rings len * weighted reference DC
    // device reference update (cra_refupdate.cu), allocated on first use
    short* d_shell = nullptr;    // [nx][nx/2+1] fsc shell of a half-complex coefficient, -1 = skipped
    int nsh = 0;                 // shells = nx/2 + 1
    std::vector<double> shell_count;   // coefficients per shell (x2: the "n" column of sp_statistics.fsc)
    double* d_fsc = nullptr;     // [max_refs][3][nsh]
    float* d_cs = nullptr;       // [max_refs][2]
};

namespace {

int ilog2_floor(int n) { int l = -1; while (n > 0) { n >>= 1; ++l; } return l; }

// Sphire Numrinit(first_ring, last_ring, skip, "F")
std::vector<int> numrinit(int ir, int ou, int rs)
{
    const int MAXFFT = 32768;
    std::vector<int> numr; int lcirc = 1;
    for (int k = ir; k <= ou; k += rs) {
        int jp = (int)(2.0 * M_PI * k + 0.5);
        int ip = 1 << (ilog2_floor(jp) + 1);
        if (k + rs <= ou && jp > ip + ip / 2) ip = std::min(MAXFFT, 2 * ip);
        if (k + rs > ou && jp > ip + ip / 5) ip = std::min(MAXFFT, 2 * ip);
        numr.push_back(k); numr.push_back(lcirc); numr.push_back(ip);
        lcirc += ip;
    }
    return numr;
}

// Phases of the grouped row kernel: consecutive 4-ring units (longest rings first) packed while
// their padded ring buffers fit the footprint of the largest unit.
int build_group_plan(CraCtx* c)
{
    const CraRingTab& t = c->htab;
    const int nring = t.nring, nunit = (nring + 3) / 4;
    auto padded = [&](int ring) { const int n = t.len[ring] >> 1, lg = ilog2_floor(n), NA = 1 << (lg / 2), NB = n / NA; return NA * (NB + 1); };
    std::vector<int> usize(nunit, 0), unk(nunit, 0);
    for (int u = 0; u < nunit; ++u) {
        for (int j = 0; j < 4; ++j) { const int ring = nring - 1 - (4 * u + j); if (ring >= 0) usize[u] += padded(ring); }
        unk[u] = t.len[nring - 1 - 4 * u] >> 1;
    }
    const int cap = *std::max_element(usize.begin(), usize.end());          // float2 per row
    // item list positions: lists run over rings nring-1 .. 0 (build_tables), the sample table over rings 0 .. nring-1
    std::vector<int> aoff(nring + 1, 0), boff(nring + 1, 0), coff(nring + 1, 0), qoff(nring + 1, 0);
    for (int s = 0; s < nring; ++s) {
        const int ring = nring - 1 - s, n = t.len[ring] >> 1, lg = ilog2_floor(n), NA = 1 << (lg / 2), NB = n / NA;
        aoff[s + 1] = aoff[s] + NB; boff[s + 1] = boff[s] + NA; coff[s + 1] = coff[s] + n / 2;
    }
    for (int i = 0; i < nring; ++i) qoff[i + 1] = qoff[i] + t.len[i] / 4;
    // rows per CTA: as many as fit two CTAs per SM (fewer image reloads, shared index math), else one CTA
    c->plan.stride = (2 * cap + 3) & ~3;
    c->plan.nring = nring;
    {   // Image tile: the whole image, or -- where that saves shared memory -- a square window around the particle's
        // search window: ring radius + taps on both sides, the window's span, and the slack of a 16-byte aligned origin.
        const int T = (2 * t.rad[nring - 1] + 2 * (int)ceilf(c->cfg.max_range) + 9 + 3) & ~3;
        c->plan.tile = (T + 8 <= c->nx) ? T : 0;
        if (const char* e = getenv("CRA_GRP_TILE")) { if (atoi(e) == 0) c->plan.tile = 0; }
    }
    int dev = 0; cudaGetDevice(&dev);
    int smem_sm = 0, smem_blk = 0;
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&smem_blk, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    auto fits = [&](int rmax, int ncta) {
        CraGroupPlan q = c->plan; q.rmax = rmax;
        const size_t need = cra_polar_group_smem(c->nx, t.maxrin, q) + 3700;   // + static shared + 1 KB reserved per CTA
        return need <= (size_t)smem_blk && need * ncta <= (size_t)smem_sm;
    };
    int rmax = 0;
    const int min2 = getenv("CRA_GRP_MIN2") ? atoi(getenv("CRA_GRP_MIN2")) : 7;      // fewest rows worth a second resident CTA
    for (int r = CRA_GRP_RMAX; r >= min2 && !rmax; --r) if (fits(r, 2)) rmax = r;
    // Two independent thread groups per CTA (cra_polar_grp.cu) where two CTAs are resident and a group keeps >= 6 rows:
    // measured 135.9 -> 125.7 ms per step at nx = 90 / ou = 36, but +11 % at nx = 128 / ou = 60, where one CTA of 17 rows
    // fills the SM and its phases hold 2 x the samples per thread.
    c->plan.nh = (rmax >= 12) ? 2 : 1;
    if (const char* e = getenv("CRA_GRP_NH")) c->plan.nh = (atoi(e) == 2) ? 2 : 1;
    for (int r = CRA_GRP_RMAX; r >= 1 && !rmax; --r) if (fits(r, 1)) rmax = r;
    c->plan.rmax = rmax;
    if (!rmax) return 0;
    const int lanes = 256 / c->plan.nh;                    // threads of one group
    std::vector<CraPhase> phases;
    std::vector<int> ppoff(nring, 0);
    auto magic = [](int n) { return n > 0 ? (1 << 24) / n + 1 : 1; };
    for (int u = 0; u < nunit;) {
        int u1 = u, used = 0;
        while (u1 < nunit && used + usize[u1] <= cap) { used += usize[u1]; ++u1; }
        const int s0 = 4 * u, s1 = std::min(4 * u1, nring);                 // slots [s0, s1)
        int acc = 0;
        for (int s = s0; s < s1; ++s) { const int ring = nring - 1 - s; ppoff[ring] = acc; acc += padded(ring); }
        CraPhase p{};
        p.u0 = u; p.u1 = u1;
        p.a0 = aoff[s0]; p.a1 = aoff[s1]; p.b0 = boff[s0]; p.b1 = boff[s1]; p.c0 = coff[s0]; p.c1 = coff[s1];
        p.q0 = qoff[nring - s1]; p.q1 = qoff[nring - s0];                   // rings nring-s1 .. nring-1-s0
        for (int uu = u; uu < u1; ++uu) p.upr += unk[uu];
        auto pick = [lanes](int n, int nr, int setup, int per_row) {
            int best = 1; long bestc = 1L << 40;
            for (int sset = 1; sset <= nr; ++sset) {
                const long cst = (long)((n * sset + lanes - 1) / lanes) * (setup + (long)((nr + sset - 1) / sset) * per_row);
                if (cst < bestc) { bestc = cst; best = sset; }
            }
            return (unsigned char)best;
        };
        for (int nr = 0; nr <= CRA_GRP_RMAX; ++nr) {
            p.nsetC[nr] = pick(p.c1 - p.c0, std::max(nr, 1), 40, 24);
            p.nsetD[nr] = pick(p.upr, std::max(nr, 1), 70, 50);
        }
        p.magicA = magic(p.a1 - p.a0); p.magicB = magic(p.b1 - p.b0); p.magicC = magic(p.c1 - p.c0); p.magicD = magic(p.upr);
        // fastdiv is exact while n * n * CRA_GRP_RMAX < 2^24
        const long worst = std::max({p.a1 - p.a0, p.b1 - p.b0, p.c1 - p.c0, p.upr});
        if (worst * worst * CRA_GRP_RMAX >= (1L << 24)) { c->plan.rmax = 0; return 0; }
        phases.push_back(p);
        u = u1;
    }
    c->plan.nphase = (int)phases.size();
    // Item order of the real-FFT split (list Cg) and of the unit gather (list D): consecutive k of one ring sit at
    // (k % NA) * (NB + 1) + k / NA in the padded ring buffer, whose shared-memory bank depends only on
    // (k % NA + k / NA) mod 16 -- a half-warp of consecutive k replays 2.4x.  Nothing ties a lane to a particular k, so
    // the items of a phase are regrouped greedily into runs of 16 whose float2 addresses fall into 16 different banks
    // (for the unit gather: of every ring slot of the unit), a host-side permutation of the work lists only.
    std::vector<int> cg_new, d_list;
    {
        auto geom = [&](int ring, int& NA, int& NB1, int& la) {
            const int n = t.len[ring] >> 1, lg = ilog2_floor(n); NA = 1 << (lg / 2); NB1 = n / NA + 1; la = lg / 2; };
        auto zpos = [&](int ring, int k) { int NA, NB1, la; geom(ring, NA, NB1, la); return ppoff[ring] + (k & (NA - 1)) * NB1 + (k >> la); };
        // greedy grouping: keys[i] = bank sets an item occupies (one per constraint); a run takes items whose banks are free
        auto regroup = [&](const std::vector<int>& items, const std::vector<std::vector<int>>& banks, std::vector<int>& out) {
            const size_t n = items.size(), nc = n ? banks[0].size() : 0;
            std::vector<char> used(n, 0);
            size_t done = 0, first = 0;
            while (done < n) {
                std::vector<unsigned> busy(nc, 0u);
                int taken = 0;
                while (first < n && used[first]) ++first;
                for (size_t i = first; i < n && taken < 16; ++i) {
                    if (used[i]) continue;
                    bool ok = true;
                    for (size_t q = 0; q < nc && ok; ++q) if (banks[i][q] >= 0 && (busy[q] >> banks[i][q]) & 1u) ok = false;
                    if (!ok) continue;
                    for (size_t q = 0; q < nc; ++q) if (banks[i][q] >= 0) busy[q] |= 1u << banks[i][q];
                    used[i] = 1; out.push_back(items[i]); ++taken; ++done;
                }
                // a run shorter than 16 is padded by what is left, conflicts accepted (the tail of a phase)
                for (size_t i = first; i < n && taken < 16 && done < n; ++i)
                    if (!used[i]) { used[i] = 1; out.push_back(items[i]); ++taken; ++done; }
            }
        };
        for (auto& ph : phases) {
            // Cg items of the phase: (ring, k), k = 1 .. n/2; banks of Z_k and of its partner Z_{n-k}
            std::vector<int> items; std::vector<std::vector<int>> banks;
            const int s0 = 4 * ph.u0, s1 = std::min(4 * ph.u1, nring);
            for (int sl = s0; sl < s1; ++sl) {
                const int ring = nring - 1 - sl, n = t.len[ring] >> 1;
                for (int k = 1; k <= n / 2; ++k) {
                    items.push_back((ring << 16) | k);
                    const int bk = zpos(ring, k) & 15, bm = zpos(ring, n - k) & 15;
                    banks.push_back({bk, (n - k == k) ? -1 : bm});
                }
            }
            if ((int)items.size() != ph.c1 - ph.c0) { cra_set_error("grouped row kernel: item list mismatch"); return 1; }
            ph.c0 = (int)cg_new.size();
            regroup(items, banks, cg_new);
            ph.c1 = (int)cg_new.size();
            // D items: (unit, k), k < the unit's longest half length; banks of the complex value each ring slot reads
            items.clear(); banks.clear();
            for (int uu = ph.u0; uu < ph.u1; ++uu)
                for (int k = 0; k < unk[uu]; ++k) {
                    items.push_back((uu << 16) | k);
                    std::vector<int> b(4, -1);
                    for (int j = 0; j < 4; ++j) {
                        const int ring = nring - 1 - (4 * uu + j);
                        if (ring < 0) continue;
                        const int n = t.len[ring] >> 1;
                        if (k > n) continue;
                        b[j] = ((k == 0 || k == n) ? ppoff[ring] : zpos(ring, k)) & 15;
                    }
                    banks.push_back(b);
                }
            ph.d0 = (int)d_list.size();
            regroup(items, banks, d_list);
            if ((int)d_list.size() - ph.d0 != ph.upr) { cra_set_error("grouped row kernel: unit list mismatch"); return 1; }
        }
    }
    const size_t bytes = sizeof(CraPhase) * phases.size() + sizeof(int) * (nring + nunit + cg_new.size() + d_list.size());
    std::vector<char> blob(bytes);
    char* w = blob.data();
    memcpy(w, phases.data(), sizeof(CraPhase) * phases.size()); w += sizeof(CraPhase) * phases.size();
    memcpy(w, ppoff.data(), sizeof(int) * nring); w += sizeof(int) * nring;
    memcpy(w, unk.data(), sizeof(int) * nunit); w += sizeof(int) * nunit;
    memcpy(w, cg_new.data(), sizeof(int) * cg_new.size()); w += sizeof(int) * cg_new.size();
    memcpy(w, d_list.data(), sizeof(int) * d_list.size());
    CRA_CUDA(cudaMalloc(&c->d_plan, bytes));
    CRA_CUDA(cudaMemcpy(c->d_plan, blob.data(), bytes, cudaMemcpyHostToDevice));
    c->plan.phases = reinterpret_cast<const CraPhase*>(c->d_plan);
    c->plan.ppoff = reinterpret_cast<const int*>(reinterpret_cast<const char*>(c->d_plan) + sizeof(CraPhase) * phases.size());
    c->plan.unit_nk = c->plan.ppoff + nring;
    c->items.Cg = c->plan.unit_nk + nunit; c->items.nCg = (int)cg_new.size();      // the regrouped lists replace build_tables' Cg
    c->items.D = c->items.Cg + cg_new.size(); c->items.nD = (int)d_list.size();
    return 0;
}

int build_tables(CraCtx* c)
{
    c->numr = numrinit(c->cfg.ir, c->cfg.ou, c->cfg.rs);
    const int nring = (int)c->numr.size() / 3;
    if (nring < 1 || nring > CRA_MAX_RINGS) { cra_set_error("ring count out of range"); return 1; }
    CraRingTab& t = c->htab;
    memset(&t, 0, sizeof(t));
    t.nring = nring;
    t.maxrin = c->numr[3 * nring - 1];
    t.lcirc = c->numr[3 * nring - 2] + c->numr[3 * nring - 1] - 1;
    t.log2n = ilog2_floor(t.maxrin);
    if ((1 << t.log2n) != t.maxrin || t.log2n < 5 || t.log2n > 10) {
        cra_set_error("maxrin must be a power of two in [32,1024] (ou between 3 and ~160)"); return 1;
    }
    float nn = 0.0f;
    std::vector<float4> samp(t.lcirc / 4);   // one quarter of every ring: x, y, ring, jt
    int pacc = 0, qacc = 0;
    std::vector<float> sampw(t.lcirc);
    const double dpi = 2 * atan(1.0);
    for (int i = 0; i < nring; ++i) {
        const int inr = c->numr[3 * i], len = c->numr[3 * i + 2], off = c->numr[3 * i + 1] - 1;
        t.off[i] = off; t.len[i] = len; t.rad[i] = inr;
        t.coff[i] = off / 2 + i;                     // sum over previous rings of (len/2 + 1)
        t.wr[i] = (float)(inr * (2.0 * M_PI) / (double)len * (double)t.maxrin / (double)len);   // ringwe
        t.wn[i] = (float)(inr * 2 * M_PI / (float)len);                                         // Normalize_ring
        const int lt = len / 4;
        const double dfi = dpi / lt;
        for (int jt = 0; jt < lt; ++jt) {
            float x, y;
            if (jt == 0) { x = 0.0f; y = (float)inr; }
            else { float fi = (float)(dfi * jt); x = sinf(fi) * inr; y = cosf(fi) * inr; }
            float4 e = make_float4(x, y, 0.f, 0.f);
            memcpy(&e.z, &i, sizeof(int)); memcpy(&e.w, &jt, sizeof(int));
            samp[qacc++] = e;
        }
        // padded smem placement: the ring is n = len/2 complex values in rows of NB (+1 pad)
        const int n = len >> 1, lgn = ilog2_floor(n), NA = 1 << (lgn / 2), NB = n / NA;
        t.poff[i] = pacc;
        pacc += NA * (NB + 1);
        for (int j = 0; j < len; ++j) { sampw[off + j] = t.wn[i]; nn += t.wn[i]; }
    }
    t.nn = nn;
    t.nc = t.lcirc / 2 + nring;
    // fragment layout: chunks of 16 rings per frequency, longest rings first (cra_common.cuh)
    {
        const int nk = t.maxrin / 2 + 1;
        c->h_koff.assign(nk + 1, 0); c->h_chunk_k.clear();
        for (int k = 0; k < nk; ++k) {
            int Kk = 0;
            for (int i = 0; i < nring; ++i) if ((t.len[i] >> 1) >= k) ++Kk;
            const int nchk = (Kk + 15) / 16;
            for (int cc = 0; cc < nchk; ++cc) c->h_chunk_k.push_back((k << 4) | cc);
            c->h_koff[k + 1] = c->h_koff[k] + nchk;
        }
        c->frag.nk = nk; c->frag.nch = c->h_koff[nk];
        std::vector<int> all(c->h_koff); all.insert(all.end(), c->h_chunk_k.begin(), c->h_chunk_k.end());
        CRA_CUDA(cudaMalloc(&c->d_fragtab, sizeof(int) * all.size()));
        CRA_CUDA(cudaMemcpy(c->d_fragtab, all.data(), sizeof(int) * all.size(), cudaMemcpyHostToDevice));
        c->frag.koff = c->d_fragtab; c->frag.chunk_k = c->d_fragtab + (nk + 1);
    }
    t.lcpad = (2 * pacc + 3) & ~3;
    std::vector<float2> twf(t.maxrin), twi;
    for (int j = 0; j < t.maxrin; ++j) {
        double a = -2.0 * M_PI * j / t.maxrin; twf[j] = make_float2((float)cos(a), (float)sin(a));
    }
    cra_ccf_twiddles(t.log2n, twi);
    // flat work lists of the ring FFT passes, longest rings first so that warps stay uniform
    {
        std::vector<int> A, B, C, Cg;
        for (int i = nring - 1; i >= 0; --i) {
            const int n = t.len[i] >> 1;
            const int lg = ilog2_floor(n), NA = 1 << (lg / 2), NB = n / NA;
            if (lg < 2 || lg > 9) { cra_set_error("ring length outside [8,1024]"); return 1; }
            for (int b = 0; b < NB; ++b) A.push_back((i << 16) | b);
            for (int a = 0; a < NA; ++a) B.push_back((i << 16) | a);
            for (int k = 0; k <= n / 2; ++k) C.push_back((i << 16) | k);
            for (int k = 1; k <= n / 2; ++k) Cg.push_back((i << 16) | k);
        }
        std::vector<int> all(A); all.insert(all.end(), B.begin(), B.end()); all.insert(all.end(), C.begin(), C.end());
        all.insert(all.end(), Cg.begin(), Cg.end());
        CRA_CUDA(cudaMalloc(&c->d_items, sizeof(int) * all.size()));
        CRA_CUDA(cudaMemcpy(c->d_items, all.data(), sizeof(int) * all.size(), cudaMemcpyHostToDevice));
        c->items.A = c->d_items; c->items.nA = (int)A.size();
        c->items.B = c->d_items + A.size(); c->items.nB = (int)B.size();
        c->items.C = c->d_items + A.size() + B.size(); c->items.nC = (int)C.size();
        c->items.Cg = c->items.C + C.size(); c->items.nCg = (int)Cg.size();
    }
    // model_circle(ou, nx, nx): r^2 <= ou^2 around (nx/2, nx/2)
    std::vector<float> mask((size_t)c->npix);
    const float rad = (float)c->cfg.ou;
    for (int j = 0; j < c->nx; ++j)
        for (int i = 0; i < c->nx; ++i) {
            float x2 = fabsf((float)i - c->nx / 2), y2 = fabsf((float)j - c->nx / 2);
            float r = (x2 * x2) / (rad * rad) + (y2 * y2) / (rad * rad);
            mask[i + (size_t)j * c->nx] = (r <= 1) ? 1.0f : 0.0f;
        }
    CRA_CUDA(cudaMalloc(&c->d_tab, sizeof(CraRingTab)));
    CRA_CUDA(cudaMemcpy(c->d_tab, &t, sizeof(CraRingTab), cudaMemcpyHostToDevice));
    CRA_CUDA(cudaMalloc(&c->d_samp, sizeof(float4) * samp.size()));
    CRA_CUDA(cudaMemcpy(c->d_samp, samp.data(), sizeof(float4) * samp.size(), cudaMemcpyHostToDevice));
    CRA_CUDA(cudaMalloc(&c->d_sampw, sizeof(float) * t.lcirc));
    CRA_CUDA(cudaMemcpy(c->d_sampw, sampw.data(), sizeof(float) * t.lcirc, cudaMemcpyHostToDevice));
    CRA_CUDA(cudaMalloc(&c->d_twf, sizeof(float2) * twf.size()));
    CRA_CUDA(cudaMemcpy(c->d_twf, twf.data(), sizeof(float2) * twf.size(), cudaMemcpyHostToDevice));
    CRA_CUDA(cudaMalloc(&c->d_twi, sizeof(float2) * twi.size()));
    CRA_CUDA(cudaMemcpy(c->d_twi, twi.data(), sizeof(float2) * twi.size(), cudaMemcpyHostToDevice));
    {
        std::vector<double2> twd(t.maxrin);
        for (int j = 0; j < t.maxrin; ++j) { const double a = 2.0 * M_PI * j / t.maxrin; twd[j] = make_double2(cos(a), sin(a)); }
        CRA_CUDA(cudaMalloc(&c->d_twd, sizeof(double2) * twd.size()));
        CRA_CUDA(cudaMemcpy(c->d_twd, twd.data(), sizeof(double2) * twd.size(), cudaMemcpyHostToDevice));
    }
    CRA_CUDA(cudaMalloc(&c->d_mask, sizeof(float) * c->npix));
    CRA_CUDA(cudaMemcpy(c->d_mask, mask.data(), sizeof(float) * c->npix, cudaMemcpyHostToDevice));
    return build_group_plan(c);
}

// dst (device) <- src (pinned host memory, read over the bus by the SMs), n 16-byte words
__global__ void pull_host_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

int window_of(const CraSearch& s, float step, int4* w)
{
    w->x = (int)(s.xl / step); w->y = (int)(s.xr / step);
    w->z = (int)(s.yl / step); w->w = (int)(s.yr / step);
    if (w->x < 0 || w->y < 0 || w->z < 0 || w->w < 0) return 1;
    return 0;
}

// order the main stream after every queued upload that overlaps particles [first, first + n)
int wait_uploads(CraCtx* c, int first, int n)
{
    for (size_t i = 0; i < c->pending.size();) {
        CraCtx::Pending& p = c->pending[i];
        if (p.first < first + n && first < p.first + p.n) {
            CRA_CUDA(cudaStreamWaitEvent(c->st, p.ev, 0));
            // normalize.mask of the whole uploaded range, once, in order with its first consumer.  On the copy stream
            // the kernel would wait for an SM slot behind the persistent CCF grid and hold back the next copy.
            if (p.sub >= 0 && cra_launch_mask_normalize(c->d_images + (size_t)p.first * c->npix, p.n, c->nx, c->d_mask, p.sub ? 0 : 2,
                                                        c->d_dc + p.first, c->st)) return 1;
            c->ev_pool.push_back(p.ev);
            c->pending.erase(c->pending.begin() + i);
        } else ++i;
    }
    return 0;
}

struct Bind { CraCtx* c; Bind(CraCtx* c_) : c(c_) {} int ok() { if (!c) { cra_set_error("null context"); return 1; }
    cudaError_t e = cudaSetDevice(c->device); if (e != cudaSuccess) { cra_set_error(cudaGetErrorString(e)); return 1; } return 0; } };

int ensure_meta(CraCtx* c, size_t nparticles, size_t nbatch_guess)
{
    size_t need = nparticles * (sizeof(CraSearch) + sizeof(int4) + 2 * sizeof(int)) + 2 * (nbatch_guess + 2) * sizeof(int) + 64;
    if (need > c->meta_bytes) {
        if (c->h_meta) cudaFreeHost(c->h_meta);
        if (c->d_meta) cudaFree(c->d_meta);
        c->h_meta = nullptr; c->d_meta = nullptr;
        c->meta_bytes = need + need / 4;
        CRA_CUDA(cudaMallocHost(&c->h_meta, c->meta_bytes));
        CRA_CUDA(cudaMalloc(&c->d_meta, c->meta_bytes));
    }
    if (nparticles > c->cap_meta) {
        if (c->d_res) cudaFree(c->d_res);
        if (c->h_res) cudaFreeHost(c->h_res);
        c->d_res = nullptr; c->h_res = nullptr;
        c->cap_meta = nparticles + nparticles / 4;
        CRA_CUDA(cudaMalloc(&c->d_res, sizeof(CraResult) * c->cap_meta));
        CRA_CUDA(cudaMallocHost(&c->h_res, sizeof(CraResult) * c->cap_meta));
    }
    return 0;
}

int ensure_par(CraCtx* c, size_t n)
{
    if (n > c->cap_par) {
        if (c->d_par) cudaFree(c->d_par);
        if (c->d_iref) cudaFree(c->d_iref);
        if (c->h_par) cudaFreeHost(c->h_par);
        if (c->h_iref) cudaFreeHost(c->h_iref);
        c->cap_par = n + n / 4;
        CRA_CUDA(cudaMalloc(&c->d_par, sizeof(float4) * c->cap_par));
        CRA_CUDA(cudaMalloc(&c->d_iref, sizeof(int) * c->cap_par));
        CRA_CUDA(cudaMallocHost(&c->h_par, sizeof(float4) * c->cap_par));
        CRA_CUDA(cudaMallocHost(&c->h_iref, sizeof(int) * c->cap_par));
    }
    return 0;
}

}  // namespace

extern "C" int cra_create(const CraConfig* cfg, int device, CraCtx** out)
{
    if (!cfg || !out) { cra_set_error("null argument"); return 1; }
    if (cfg->nx < 8 || cfg->ir < 1 || cfg->ou < cfg->ir || cfg->rs < 1 || cfg->step <= 0.f ||
        cfg->max_particles < 1 || cfg->max_refs < 1 || cfg->max_range < 0.f) {
        cra_set_error("invalid CraConfig"); return 1;
    }
    if (cfg->ou + 2 > cfg->nx / 2) { cra_set_error("ou too large for nx (needs ou <= nx/2 - 2)"); return 1; }
    int ndev = 0;
    CRA_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { cra_set_error("no such CUDA device"); return 1; }
    CRA_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CRA_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { cra_set_error("this engine is built for sm_100a (B200) only"); return 1; }
    CraCtx* c = new CraCtx();
    c->cfg = *cfg; c->device = device; c->nx = cfg->nx; c->npix = cfg->nx * cfg->nx;
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) { cra_set_error("stream create failed"); delete c; return 1; }
    {
        const char* e = getenv("CRA_CCF");
        if (e && strcmp(e, "simt") == 0) c->fmt = CRA_FMT_F32;
        else if (e && strcmp(e, "mma") == 0) c->use_tm = false;
        else if (e && strcmp(e, "um") == 0) c->use_um = true;
        else if (e && strcmp(e, "tm") != 0 && e[0]) { cra_set_error("CRA_CCF must be 'um', 'tm', 'mma' or 'simt'"); cra_destroy(c); return 1; }
        const char* pk = getenv("CRA_POLAR");
        if (pk && strcmp(pk, "general") == 0) c->use_group = false;
        else if (pk && strcmp(pk, "group") != 0 && pk[0]) { cra_set_error("CRA_POLAR must be 'group' or 'general'"); cra_destroy(c); return 1; }
    }
    if (build_tables(c)) { cra_destroy(c); return 1; }
    {   // the references (and steps / centres the grouped kernel does not cover) go through the general row kernel:
        // whole polar rows in shared memory, and the image beside them when it fits (boxes up to ~170 pixels)
        int smem_blk = 0;
        cudaDeviceGetAttribute(&smem_blk, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
        int one = 1;
        if (cra_polar_general_layout(c->nx, c->htab, (size_t)smem_blk, cra_polar_default_rpb(), &c->rpb, &c->gimg_rows) ||
            cra_polar_general_layout(c->nx, c->htab, (size_t)smem_blk, 1, &one, &c->gimg_one)) {
            cra_set_error("ou too large for this build: one polar row (" + std::to_string((size_t)c->htab.lcpad * 4 / 1024) +
                          " KB) does not fit into shared memory (limit " + std::to_string(smem_blk / 1024) + " KB); reduce ou");
            cra_destroy(c); return 1;
        }
    }
    const int k = (int)(cfg->max_range / cfg->step);
    c->smax = (2 * k + 1) * (2 * k + 1);
    const size_t row_bytes = (c->fmt == CRA_FMT_FRAG) ? cra_frag_row_bytes(c->frag.nch)
                                                      : (size_t)c->htab.nc * sizeof(float2);     // device spectrum of one row
    c->row_bytes = row_bytes;
    long rb = cfg->row_batch > 0 ? cfg->row_batch : (long)((size_t)2 << 30) / (long)row_bytes;
    if (rb < c->smax) rb = c->smax;
    long want = (long)cfg->max_particles * c->smax;
    if (rb > want) rb = want;
    c->row_batch = (int)rb;
    if (!cra_ccf_tm_supported(c->htab.log2n)) c->use_tm = false;
    if (c->fmt != CRA_FMT_FRAG || !cra_ccf_um_supported(c->htab.log2n)) c->use_um = false;
    c->frag.unit_rows = c->use_um ? 1 : 0;
    c->ntile_n_max = (c->fmt == CRA_FMT_FRAG) ? (c->use_um ? cra_ccf_um_num_tiles(cfg->max_refs)
                                                 : c->use_tm ? cra_ccf_tm_num_tiles(cfg->max_refs, c->htab.log2n)
                                                           : cra_ccf_mma_num_tiles(cfg->max_refs, c->htab.log2n))
                                              : (cfg->max_refs + cra_ccf_tile_n() - 1) / cra_ccf_tile_n();
    const size_t nsum = (size_t)cfg->max_refs * 2 * c->npix + cfg->max_refs;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&c->d_images, (size_t)cfg->max_particles * c->npix * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_dc, (size_t)cfg->max_particles * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(c->d_dc, 0, (size_t)cfg->max_particles * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_refs, (size_t)cfg->max_refs * c->npix * sizeof(float));
    const size_t ref_groups = ((size_t)cfg->max_refs + 3) / 4, row_groups = ((size_t)c->row_batch + 3) / 4;
    // one extra quad: a class-bound launch (cra_align_bound) bases the 4-reference operand loads at any reference
    if (e == cudaSuccess) e = cudaMalloc(&c->d_refspec, (ref_groups + 1) * 4 * row_bytes);
    if (e == cudaSuccess) e = cudaMemset(c->d_refspec, 0, (ref_groups + 1) * 4 * row_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_spec, row_groups * 4 * row_bytes);
    if (e == cudaSuccess) e = cudaMemset(c->d_spec, 0, row_groups * 4 * row_bytes);
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_group, 4 * row_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_cand, (size_t)c->row_batch * c->ntile_n_max * sizeof(CraCand));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_norm, ((size_t)c->row_batch + 4) * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_tref, (ref_groups + 1) * 4 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(c->d_tref, 0, (ref_groups + 1) * 4 * sizeof(float));
    if (e == cudaSuccess && c->use_um) e = cudaMalloc(&c->d_refimg, cra_ccf_um_refimg_bytes(cfg->max_refs, c->frag.nch));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_sums, nsum * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_curves, (size_t)2 * c->htab.maxrin * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(c->d_sums, 0, nsum * sizeof(float));
    if (e != cudaSuccess) { cra_set_error(std::string("device allocation failed: ") + cudaGetErrorString(e)); cra_destroy(c); return 1; }
    *out = c;
    return 0;
}

extern "C" int cra_destroy(CraCtx* c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->st_copy) cudaStreamSynchronize(c->st_copy);
    if (c->st2) cudaStreamSynchronize(c->st2);
    if (c->st) cudaStreamSynchronize(c->st);
    for (auto& e : c->ev) cudaEventDestroy(e);
    for (auto& p : c->pending) cudaEventDestroy(p.ev);
    for (auto& e : c->ev_pool) cudaEventDestroy(e);
    if (c->st_copy) cudaStreamDestroy(c->st_copy);
    cra_ccf_tm_sched_free(c->tm_sched); c->tm_sched = nullptr;
    cudaFree(c->d_items); cudaFree(c->d_fragtab); cudaFree(c->d_norm); cudaFree(c->d_tref); cudaFree(c->d_plan); cudaFree(c->d_refimg);
    cudaFree(c->d_tab); cudaFree(c->d_samp); cudaFree(c->d_sampw); cudaFree(c->d_twf); cudaFree(c->d_twi); cudaFree(c->d_twd);
    cudaFree(c->d_mask); cudaFree(c->d_dc); cudaFree(c->d_images); cudaFree(c->d_refs); cudaFree(c->d_refspec);
    cudaFree(c->d_spec); cudaFree(c->d_cand); cudaFree(c->d_sums); cudaFree(c->d_meta); cudaFree(c->d_res);
    cudaFree(c->d_spec2); cudaFree(c->d_cand2); cudaFree(c->d_norm2);
    for (auto& e : c->ev_lane) if (e) cudaEventDestroy(e);
    if (c->st2) cudaStreamDestroy(c->st2);
    cudaFree(c->d_par); cudaFree(c->d_iref); cudaFree(c->d_tmpimg); cudaFree(c->d_curves);
    cudaFree(c->d_shell); cudaFree(c->d_fsc); cudaFree(c->d_cs); cudaFree(c->d_dft);
    if (c->h_meta) cudaFreeHost(c->h_meta);
    if (c->h_res) cudaFreeHost(c->h_res);
    if (c->h_par) cudaFreeHost(c->h_par);
    if (c->h_iref) cudaFreeHost(c->h_iref);
    if (c->h_group) cudaFreeHost(c->h_group);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
    return 0;
}

extern "C" int cra_ring_info(CraCtx* c, int* nring, int* lcirc, int* maxrin, int* numr_out)
{
    if (!c) { cra_set_error("null context"); return 1; }
    if (nring) *nring = c->htab.nring;
    if (lcirc) *lcirc = c->htab.lcirc;
    if (maxrin) *maxrin = c->htab.maxrin;
    if (numr_out) memcpy(numr_out, c->numr.data(), c->numr.size() * sizeof(int));
    return 0;
}

static int upload_particles(CraCtx* c, const float* src, int first, int n, int sub, cudaMemcpyKind kind)
{
    CraNvtx range("cra_upload_particles");
    Bind b(c); if (b.ok()) return 1;
    if (first < 0 || n < 0 || first + n > c->cfg.max_particles) { cra_set_error("particle range exceeds max_particles"); return 1; }
    if (n == 0) return 0;
    if (wait_uploads(c, first, n)) return 1;
    float* dst = c->d_images + (size_t)first * c->npix;
    CRA_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * c->npix * sizeof(float), kind, c->st));
    if (cra_launch_mask_normalize(dst, n, c->nx, c->d_mask, sub ? 0 : 2, c->d_dc + first, c->st)) return 1;
    CRA_CUDA(cudaStreamSynchronize(c->st));
    return 0;
}
extern "C" int cra_upload_particles(CraCtx* c, const float* h, int first, int n, int sub)
{ return upload_particles(c, h, first, n, sub, cudaMemcpyHostToDevice); }

extern "C" int cra_upload_particles_async(CraCtx* c, const float* h, int first, int n, int sub)
{
    CraNvtx range("cra_upload_particles_async");
    Bind b(c); if (b.ok()) return 1;
    if (first < 0 || n < 0 || first + n > c->cfg.max_particles) { cra_set_error("particle range exceeds max_particles"); return 1; }
    if (n == 0) return 0;
    if (!c->st_copy) CRA_CUDA(cudaStreamCreateWithFlags(&c->st_copy, cudaStreamNonBlocking));
    // the slots may still be read by work already queued on the main stream
    cudaEvent_t e0;
    if (c->ev_pool.empty()) CRA_CUDA(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming)); else { e0 = c->ev_pool.back(); c->ev_pool.pop_back(); }
    CRA_CUDA(cudaEventRecord(e0, c->st));
    CRA_CUDA(cudaStreamWaitEvent(c->st_copy, e0, 0));
    float* dst = c->d_images + (size_t)first * c->npix;
    CRA_CUDA(cudaMemcpyAsync(dst, h, (size_t)n * c->npix * sizeof(float), cudaMemcpyHostToDevice, c->st_copy));
    CRA_CUDA(cudaEventRecord(e0, c->st_copy));
    // an older pending upload that this one overwrites completely no longer needs its mask subtraction
    for (auto& p : c->pending) if (p.first >= first && p.first + p.n <= first + n) p.sub = -1;
    c->pending.push_back({first, n, sub ? 1 : 0, e0});
    return 0;
}

extern "C" int cra_upload_wait(CraCtx* c)
{
    Bind b(c); if (b.ok()) return 1;
    if (c->st_copy) CRA_CUDA(cudaStreamSynchronize(c->st_copy));
    if (wait_uploads(c, 0, c->cfg.max_particles)) return 1;         // pending mask subtractions
    CRA_CUDA(cudaStreamSynchronize(c->st));
    return 0;
}
extern "C" int cra_upload_particles_dev(CraCtx* c, const float* d, int first, int n, int sub)
{ return upload_particles(c, d, first, n, sub, cudaMemcpyDeviceToDevice); }

// d_refs[0..R) -> (normalize.mask no_sigma=1) -> Polar2Dm + Frngs + Applyws -> refspec, tref
// global scratch of the reference-update transforms, only for boxes whose spectra exceed shared memory
static int ensure_dft_scratch(CraCtx* c, int n, int nsh, int nspec)
{
    const size_t per = cra_dft_scratch_elems(c->nx, nsh, nspec);
    if (!per) return 0;
    const size_t need = per * (size_t)n;
    if (need <= c->dft_elems) return 0;
    CRA_CUDA(cudaStreamSynchronize(c->st));
    cudaFree(c->d_dft); c->d_dft = nullptr; c->dft_elems = 0;
    if (cudaMalloc(&c->d_dft, need * sizeof(float2)) != cudaSuccess) { cra_set_error("device allocation failed: reference-update scratch"); return 1; }
    c->dft_elems = need;
    return 0;
}

static int prepare_refs(CraCtx* c, int R, int normalize_mask)
{
    CraNvtx range("cra_prepare_refs");
    if (normalize_mask && cra_launch_mask_normalize(c->d_refs, R, c->nx, c->d_mask, 1, nullptr, c->st)) return 1;
    if (cra_launch_polar_refs(c->d_refs, R, c->nx, c->d_tab, c->htab, c->d_samp, c->d_twf, c->items, c->d_refspec,
                              c->fmt, c->frag, c->d_tref, c->gimg_one, c->st)) return 1;
    if (c->use_um && cra_ccf_um_pack_refs(reinterpret_cast<const unsigned char*>(c->d_refspec), R, c->frag, c->d_refimg, c->st)) return 1;
    CRA_CUDA(cudaStreamSynchronize(c->st));
    c->R = R;
    return 0;
}

extern "C" int cra_set_refs(CraCtx* c, const float* h, int R, int normalize_mask)
{
    Bind b(c); if (b.ok()) return 1;
    if (R < 1 || R > c->cfg.max_refs) { cra_set_error("R exceeds max_refs"); return 1; }
    CRA_CUDA(cudaMemcpyAsync(c->d_refs, h, (size_t)R * c->npix * sizeof(float), cudaMemcpyHostToDevice, c->st));
    return prepare_refs(c, R, normalize_mask);
}

extern "C" int cra_refs_from_sums(CraCtx* c, int normalize_mask)
{
    Bind b(c); if (b.ok()) return 1;
    if (c->R < 1) { cra_set_error("cra_set_refs has not been called"); return 1; }
    const float* counts = c->d_sums + (size_t)c->cfg.max_refs * 2 * c->npix;
    if (cra_launch_class_average(c->d_sums, counts, c->d_refs, c->R, c->nx, c->st)) return 1;
    return prepare_refs(c, c->R, normalize_mask);
}

extern "C" int cra_filter_refs(CraCtx* c, float cutoff, float falloff, int normalize_mask)
{
    Bind b(c); if (b.ok()) return 1;
    if (c->R < 1) { cra_set_error("cra_set_refs has not been called"); return 1; }
    if (ensure_dft_scratch(c, c->R, 0, 2)) return 1;
    if (cra_launch_tanl_filter(c->d_refs, c->R, c->nx, cutoff, falloff, c->d_dft, c->st)) return 1;
    return prepare_refs(c, c->R, normalize_mask);
}

extern "C" int cra_get_refs(CraCtx* c, float* host_refs)
{
    Bind b(c); if (b.ok()) return 1;
    if (c->R < 1 || !host_refs) { cra_set_error("no references / null buffer"); return 1; }
    CRA_CUDA(cudaMemcpyAsync(host_refs, c->d_refs, (size_t)c->R * c->npix * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    CRA_CUDA(cudaStreamSynchronize(c->st));
    return 0;
}

// ---- reference update on the device (cra_refupdate.cu) -------------------------------------------------
// fsc's shell of every half-complex coefficient, exactly as EMData::calc_fourier_shell_correlation bins them
// (float32 argument of the square root, round half away from zero; Friedel mates on the kx = 0 column skipped)
static int ensure_fsc_tables(CraCtx* c)
{
    if (c->d_shell) return 0;
    const int nx = c->nx, nh = nx / 2 + 1, nx2 = nx / 2;
    const int inc = (int)(nx2 + 0.5);
    c->nsh = inc + 1;
    std::vector<short> sh((size_t)nx * nh);
    c->shell_count.assign(c->nsh, 0.0);
    for (int iy = 0; iy < nx; ++iy) {
        const int ky = iy > nx / 2 ? iy - nx : iy;
        for (int kx = 0; kx < nh; ++kx) {
            const float arg = (float)((double)(ky * ky) / (double)(nx2 * nx2) + (double)(kx * kx) / (double)(nx2 * nx2));
            const float argx = 0.5f * sqrtf(arg);
            const double v = (double)inc * 2.0 * (double)argx;
            const long r = (long)(v >= 0 ? floor(v + 0.5) : ceil(v - 0.5));
            const bool use = (kx > 0 || ky >= 0) && r <= inc;
            sh[(size_t)iy * nh + kx] = use ? (short)r : (short)-1;
            if (use) c->shell_count[r] += 1.0;
        }
    }
    CRA_CUDA(cudaMalloc(&c->d_shell, sh.size() * sizeof(short)));
    CRA_CUDA(cudaMemcpy(c->d_shell, sh.data(), sh.size() * sizeof(short), cudaMemcpyHostToDevice));
    CRA_CUDA(cudaMalloc(&c->d_fsc, (size_t)c->cfg.max_refs * 3 * c->nsh * sizeof(double)));
    CRA_CUDA(cudaMalloc(&c->d_cs, (size_t)c->cfg.max_refs * 2 * sizeof(float)));
    return 0;
}

extern "C" int cra_class_fsc(CraCtx* c, int masked, int min_members, int write_avg, float avg_div, int* nshell,
                             double* freq_out, double* fsc_out, double* n_out, float* counts_out)
{
    CraNvtx range("cra_class_fsc");
    Bind b(c); if (b.ok()) return 1;
    if (c->R < 1) { cra_set_error("cra_set_refs has not been called"); return 1; }
    if (ensure_fsc_tables(c)) return 1;
    const int R = c->R, nsh = c->nsh;
    if (nshell) *nshell = nsh;
    if (!fsc_out) return 0;                                    // size query
    const float* counts = c->d_sums + (size_t)c->cfg.max_refs * 2 * c->npix;
    if (ensure_dft_scratch(c, R, nsh, 3)) return 1;
    if (cra_launch_class_fsc(c->d_sums, counts, c->d_refs, c->d_shell, c->d_mask, R, c->nx, nsh, masked, min_members,
                             write_avg, avg_div, c->d_fsc, c->d_dft, c->st)) return 1;
    std::vector<double> h((size_t)R * 3 * nsh);
    std::vector<float> hc(R);
    CRA_CUDA(cudaMemcpyAsync(h.data(), c->d_fsc, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CRA_CUDA(cudaMemcpyAsync(hc.data(), counts, R * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    CRA_CUDA(cudaStreamSynchronize(c->st));
    for (int r = 0; r < R; ++r) {
        if (counts_out) counts_out[r] = hc[r];
        for (int i = 0; i < nsh; ++i) {
            const double num = h[((size_t)r * 3 + 0) * nsh + i], n1 = h[((size_t)r * 3 + 1) * nsh + i], n2 = h[((size_t)r * 3 + 2) * nsh + i];
            const double den = sqrt(n1 * n2);
            // the curve is a float in EMAN2 (calc_fourier_shell_correlation returns vector<float>)
            fsc_out[(size_t)r * nsh + i] = (hc[r] >= (float)min_members && den > 0.0) ? (double)(float)(num / den) : 0.0;
        }
    }
    for (int i = 0; i < nsh; ++i) {
        if (freq_out) freq_out[i] = (double)i / (double)(2 * (nsh - 1));
        if (n_out) n_out[i] = 2.0 * c->shell_count[i];
    }
    return 0;
}

extern "C" int cra_put_ref(CraCtx* c, int iref, const float* host_img)
{
    Bind b(c); if (b.ok()) return 1;
    if (iref < 0 || iref >= c->R || !host_img) { cra_set_error("cra_put_ref: reference index outside the current set"); return 1; }
    CRA_CUDA(cudaMemcpyAsync(c->d_refs + (size_t)iref * c->npix, host_img, (size_t)c->npix * sizeof(float), cudaMemcpyHostToDevice, c->st));
    CRA_CUDA(cudaStreamSynchronize(c->st));
    return 0;
}

extern "C" int cra_filter_center_refs(CraCtx* c, float cutoff, float falloff, int mode, float sx, float sy,
                                      int normalize_mask, float* cs_out)
{
    CraNvtx range("cra_filter_center_refs");
    Bind b(c); if (b.ok()) return 1;
    if (c->R < 1) { cra_set_error("cra_set_refs has not been called"); return 1; }
    if (mode < 0 || mode > 2) { cra_set_error("cra_filter_center_refs: mode must be 0 (filter), 1 (phase centre) or 2 (given shift)"); return 1; }
    if (ensure_fsc_tables(c)) return 1;
    if (ensure_dft_scratch(c, c->R, 0, 2)) return 1;
    if (cra_launch_filter_center(c->d_refs, c->R, c->nx, cutoff, falloff, mode, sx, sy, c->d_cs, c->d_dft, c->st)) return 1;
    if (normalize_mask && cra_launch_mask_normalize(c->d_refs, c->R, c->nx, c->d_mask, 1, nullptr, c->st)) return 1;
    if (cs_out) CRA_CUDA(cudaMemcpyAsync(cs_out, c->d_cs, (size_t)c->R * 2 * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    CRA_CUDA(cudaStreamSynchronize(c->st));
    return 0;
}

extern "C" int cra_prepare_refs(CraCtx* c, int normalize_mask)
{
    Bind b(c); if (b.ok()) return 1;
    if (c->R < 1) { cra_set_error("no references on the device"); return 1; }
    return prepare_refs(c, c->R, normalize_mask);
}

// class_of == nullptr: every particle against every reference (Util.multiref_polar_ali_2d).
// class_of != nullptr: particle p against reference class_of[p] only (gpu_isac's class-bound
// ref_free_alignment_2D, cuda/gpu_aln_noref.cu:743-782): the row kernel runs per batch as usual,
// the CCF and finalize kernels run once per run of consecutive particles of one class, on that
// run's rows and with the reference operands based at that class.
// The second pipeline lane, built on first use (2 GB of spectra at the default row batch)
static int ensure_lane2(CraCtx* c)
{
    if (c->st2) return 0;
    const size_t row_groups = ((size_t)c->row_batch + 3) / 4;
    cudaError_t e = cudaStreamCreateWithFlags(&c->st2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_spec2, row_groups * 4 * c->row_bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_spec2, 0, row_groups * 4 * c->row_bytes, c->st2);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_cand2, (size_t)c->row_batch * c->ntile_n_max * sizeof(CraCand));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_norm2, ((size_t)c->row_batch + 4) * sizeof(float2));
    for (auto& ev : c->ev_lane) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->st2);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(c->d_spec2); cudaFree(c->d_cand2); cudaFree(c->d_norm2); c->d_spec2 = nullptr; c->d_cand2 = nullptr; c->d_norm2 = nullptr;
        for (auto& ev : c->ev_lane) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
        if (c->st2) { cudaStreamDestroy(c->st2); c->st2 = nullptr; }
        return 1;                                   // not an error of the call: the single-lane path serves it
    }
    return 0;
}

static int align_impl(CraCtx* c, int start, int stop, const CraSearch* search, const int* class_of, CraResult* out)
{
    CraNvtx range(class_of ? "cra_align_bound" : "cra_align");
    Bind b(c); if (b.ok()) return 1;
    const int n = stop - start;
    if (start < 0 || n < 0 || stop > c->cfg.max_particles) { cra_set_error("particle range exceeds max_particles"); return 1; }
    if (c->R < 1) { cra_set_error("cra_set_refs has not been called"); return 1; }
    if (n == 0) return 0;
    if (class_of) {
        if (c->fmt != CRA_FMT_FRAG || c->use_um) { cra_set_error("class-bound alignment needs the fragment layout with the tm / mma kernels (default)"); return 1; }
        for (int p = 0; p < n; ++p)
            if (class_of[p] < 0 || class_of[p] >= c->R) { cra_set_error("class_of entry outside the reference list"); return 1; }
    }
    if (wait_uploads(c, start, n)) return 1;
    const float step = c->cfg.step;
    // plan batches
    std::vector<int> bfirst, bcount, brows;
    {
        int cur_first = 0, cur_rows = 0;
        for (int p = 0; p < n; ++p) {
            int4 w;
            if (window_of(search[p], step, &w)) { cra_set_error("negative search window"); return 1; }
            const long rows = (long)(w.x + w.y + 1) * (w.z + w.w + 1);
            if (rows > c->row_batch) { cra_set_error("search window larger than max_range allows"); return 1; }
            if (cur_rows + rows > c->row_batch) {
                bfirst.push_back(cur_first); bcount.push_back(p - cur_first); brows.push_back(cur_rows);
                cur_first = p; cur_rows = 0;
            }
            cur_rows += (int)rows;
        }
        bfirst.push_back(cur_first); bcount.push_back(n - cur_first); brows.push_back(cur_rows);
    }
    const size_t nb = bfirst.size();
    if (ensure_meta(c, n, nb)) return 1;
    // meta layout: win[n] (16-byte aligned) | search[n] | row_start[n + nb] | chunk_start[n + nb]
    int4* h_win = reinterpret_cast<int4*>(c->h_meta);
    CraSearch* h_search = reinterpret_cast<CraSearch*>(c->h_meta + (size_t)n * sizeof(int4));
    int* h_rs = reinterpret_cast<int*>(c->h_meta + (size_t)n * (sizeof(CraSearch) + sizeof(int4)));
    int* h_cs = h_rs + n + nb;
    const int rpb = c->rpb;
    std::vector<int> bchunks(nb);
    std::vector<char> bgroup(nb, 0);
    memcpy(h_search, search, (size_t)n * sizeof(CraSearch));
    long total_rows = 0;
    // the grouped row kernel needs a step its phase classes cover and every sample in [2, nx] (1-based), which is
    // what search_range guarantees: its tile is the bare image, and only a sample exactly on the last column / row
    // (a pixel boundary: redone with quadri's circular closure) may touch the wrap-around neighbour
    const int sub = cra_group_sub(step);                       // phase classes per axis (0: the general kernel)
    const bool group_cfg = c->fmt == CRA_FMT_FRAG && c->use_group && c->plan.rmax > 0 && sub > 0;
    // (a windowed tile holds no wrap-around neighbours: its samples must also stay off the last column / row)
    const float rmaxf = (float)c->htab.rad[c->htab.nring - 1], lo_ok = 2.0f + rmaxf, hi_ok = (float)c->nx - rmaxf - (c->plan.tile ? 1.0f : 0.0f);
    for (size_t bi = 0; bi < nb; ++bi) {
        int* rs = h_rs + bfirst[bi] + bi;
        int* cs = h_cs + bfirst[bi] + bi;
        bool grp = group_cfg;
        for (int q = 0; q < bcount[bi] && grp; ++q) {
            const int p = bfirst[bi] + q;
            int4 w; window_of(search[p], step, &w);
            const CraSearch& sp = search[p];
            grp = (sp.cx - w.x * step >= lo_ok) && (sp.cx + w.y * step <= hi_ok) &&
                  (sp.cy - w.z * step >= lo_ok) && (sp.cy + w.w * step <= hi_ok);
            // a windowed tile is sized for windows of up to 2 ceil(max_range) pixels: a wider request takes the general kernel
            if (c->plan.tile && (float)std::max(w.x + w.y, w.z + w.w) * step > 2.0f * ceilf(c->cfg.max_range)) grp = false;
        }
        bgroup[bi] = grp;
        int acc = 0, cacc = 0;
        for (int q = 0; q < bcount[bi]; ++q) {
            const int p = bfirst[bi] + q;
            window_of(search[p], step, &h_win[p]);
            rs[q] = acc; cs[q] = cacc;
            const int rows = (h_win[p].x + h_win[p].y + 1) * (h_win[p].z + h_win[p].w + 1);
            if (grp) {                                                    // balanced row blocks of <= rmax rows per phase class
                const int wx = h_win[p].x + h_win[p].y + 1, wy = h_win[p].z + h_win[p].w + 1;
                for (int cy = 0; cy < sub; ++cy)
                    for (int cx = 0; cx < sub; ++cx) {
                        const int rc = ((wx - cx + sub - 1) / sub) * ((wy - cy + sub - 1) / sub);
                        cacc += (rc + c->plan.rmax - 1) / c->plan.rmax;
                    }
            }
            else cacc += (acc + rows - 1) / rpb - acc / rpb + 1;          // aligned sub-groups this particle touches
            acc += rows;
        }
        rs[bcount[bi]] = acc; cs[bcount[bi]] = cacc;
        bchunks[bi] = cacc;
        total_rows += acc;
    }
    const size_t used = (size_t)n * (sizeof(CraSearch) + sizeof(int4)) + 2 * (n + nb) * sizeof(int);
    // The request tables go to the device through a kernel that reads the pinned host buffer directly, not
    // through the copy engine: a cudaMemcpyAsync would queue behind every particle upload already submitted
    // (cra_upload_particles_async) and hold the first alignment back by the whole stack's transfer time.
    pull_host_kernel<<<(unsigned)((used / 16 + 1 + 255) / 256), 256, 0, c->st>>>(
        reinterpret_cast<uint4*>(c->d_meta), reinterpret_cast<const uint4*>(c->h_meta), used / 16 + 1);
    CRA_CUDA(cudaGetLastError());
    const int4* d_win = reinterpret_cast<const int4*>(c->d_meta);
    const CraSearch* d_search = reinterpret_cast<const CraSearch*>(c->d_meta + (size_t)n * sizeof(int4));
    const int* d_rs = reinterpret_cast<const int*>(c->d_meta + (size_t)n * (sizeof(CraSearch) + sizeof(int4)));
    const int* d_cs = d_rs + n + nb;

    const int TN = cra_ccf_tile_n();
    const int ntile_n = (c->fmt == CRA_FMT_FRAG) ? (c->use_um ? cra_ccf_um_num_tiles(c->R)
                                                    : c->use_tm ? cra_ccf_tm_num_tiles(c->R, c->htab.log2n)
                                                              : cra_ccf_mma_num_tiles(c->R, c->htab.log2n)) : (c->R + TN - 1) / TN;
    const bool tm = c->timing;
    if (tm) {
        while (c->ev.size() < 4 * nb) { cudaEvent_t e; CRA_CUDA(cudaEventCreate(&e)); c->ev.push_back(e); }
    }
    long launches = 0;
    // Two pipeline lanes (CRA_LANES=2; measured, NOT the default): odd batches run on a second stream with their own
    // spectrum / candidate / norm buffers, so that the CTAs of one batch's row kernel fill the SM slots the other batch's
    // persistent CCF grid frees at its tail (and the reverse); kernels of one lane stay in order on their stream.
    // Identical results, but 55.2 against 52.6 ms per 16 384 particles of configuration 2 (scripts/lane_probe.py): a row
    // CTA and a CCF CTA sharing an SM run slower than two CTAs of one kernel, which costs more than the kernel tails
    // (~3 % each) give back.  Never with per-kernel timing or the class-bound variant.
    static const bool lanes_env = getenv("CRA_LANES") && atoi(getenv("CRA_LANES")) == 2;
    const bool two_lanes = lanes_env && !tm && !class_of && nb >= 2 && c->fmt == CRA_FMT_FRAG && !c->use_um && ensure_lane2(c) == 0;
    if (two_lanes) {                                 // lane 2 starts after everything queued on the main stream so far
        CRA_CUDA(cudaEventRecord(c->ev_lane[0], c->st));
        CRA_CUDA(cudaStreamWaitEvent(c->st2, c->ev_lane[0], 0));
    }
    for (size_t bi = 0; bi < nb; ++bi) {
        const bool lane2 = two_lanes && (bi & 1);
        struct Lane { cudaStream_t st; float* spec; CraCand* cand; float2* norm; };
        const Lane L = lane2 ? Lane{c->st2, c->d_spec2, c->d_cand2, c->d_norm2} : Lane{c->st, c->d_spec, c->d_cand, c->d_norm};
        CraRowMap map;
        map.row_start = d_rs + bfirst[bi] + bi;
        map.chunk_start = d_cs + bfirst[bi] + bi; map.nchunks = bchunks[bi];
        map.search = d_search + bfirst[bi];
        map.win = d_win + bfirst[bi];
        map.np = bcount[bi]; map.nrows = brows[bi]; map.p0 = start + bfirst[bi]; map.step = step; map.dc = c->d_dc;
        if (tm) CRA_CUDA(cudaEventRecord(c->ev[4 * bi + 0], L.st));
        nvtxRangePushA("row kernel (Polar2Dm + Frngs)");
        if (bgroup[bi]) {
            if (cra_launch_polar_group(c->d_images, c->nx, c->d_tab, c->htab, c->d_samp, c->d_twf, c->items, c->plan, map,
                                       c->cfg.normalize_ring, L.spec, c->frag, L.norm, L.st)) { nvtxRangePop(); return 1; }
        } else if (cra_launch_polar_rows(c->d_images, c->nx, c->d_tab, c->htab, c->d_samp, c->d_sampw, c->d_twf, c->items, map,
                                         c->cfg.normalize_ring, L.spec, c->fmt, c->frag, L.norm, c->rpb, c->gimg_rows, L.st)) { nvtxRangePop(); return 1; }
        nvtxRangePop();
        if (tm) CRA_CUDA(cudaEventRecord(c->ev[4 * bi + 1], L.st));
        if (class_of) {
            const unsigned char* specb = reinterpret_cast<const unsigned char*>(L.spec);
            const unsigned char* refb = reinterpret_cast<const unsigned char*>(c->d_refspec);
            const int* rs = h_rs + bfirst[bi] + bi;                 // batch-local first row of each particle
            for (int a = 0; a < bcount[bi];) {
                const int cls = class_of[bfirst[bi] + a];
                int e = a + 1;
                while (e < bcount[bi] && class_of[bfirst[bi] + e] == cls) ++e;
                const int r0 = rs[a], nr = rs[e] - rs[a];
                const unsigned char* ref1 = refb + (size_t)cls * c->row_bytes;
                if (nr > 0) {
                    if (c->use_tm) {
                        if (cra_launch_ccf_tm(specb + (size_t)r0 * c->row_bytes, nr, ref1, 1, c->htab, c->frag, c->h_koff, c->d_twi,
                                              L.cand + r0, 1, L.norm + r0, c->d_tref + cls, &c->tm_sched, L.st)) return 1;
                    } else if (cra_launch_ccf_mma(specb + (size_t)r0 * c->row_bytes, nr, ref1, 1, c->htab, c->frag, c->h_koff, c->d_twi,
                                                  L.cand + r0, 1, L.norm + r0, c->d_tref + cls, L.st)) return 1;
                }
                CraRowMap sub = map;
                sub.row_start += a; sub.search += a; sub.win += a; sub.np = e - a; sub.p0 += a;
                if (cra_launch_finalize(L.spec, reinterpret_cast<const float*>(ref1), 1, c->d_tab, c->htab, L.cand, 1, sub,
                                        c->d_res + bfirst[bi] + a, c->fmt, c->frag, c->d_twd, L.norm, c->d_tref + cls, L.st)) return 1;
                launches += 2;
                a = e;
            }
            if (tm) { CRA_CUDA(cudaEventRecord(c->ev[4 * bi + 2], L.st)); CRA_CUDA(cudaEventRecord(c->ev[4 * bi + 3], L.st)); }
            launches += 1;
            c->last_rows = map.nrows; c->last_group = bgroup[bi]; c->last_spec = L.spec; c->last_norm = L.norm;
            continue;
        }
        CraNvtx ccf_range("CCF (Crosrng_ms + inverse FFT + peak) + finalize");
        if (c->fmt == CRA_FMT_FRAG && c->use_um) {
            if (cra_launch_ccf_um(reinterpret_cast<const unsigned char*>(L.spec), map.nrows, c->d_refimg, c->R, c->htab, c->frag,
                                  c->h_koff, c->d_twi, L.cand, ntile_n, L.norm, c->d_tref, L.st)) return 1;
        } else if (c->fmt == CRA_FMT_FRAG && c->use_tm) {
            if (cra_launch_ccf_tm(reinterpret_cast<const unsigned char*>(L.spec), map.nrows,
                                  reinterpret_cast<const unsigned char*>(c->d_refspec), c->R, c->htab, c->frag, c->h_koff,
                                  c->d_twi, L.cand, ntile_n, L.norm, c->d_tref, &c->tm_sched, L.st)) return 1;
        } else if (c->fmt == CRA_FMT_FRAG) {
            if (cra_launch_ccf_mma(reinterpret_cast<const unsigned char*>(L.spec), map.nrows,
                                   reinterpret_cast<const unsigned char*>(c->d_refspec), c->R, c->htab, c->frag, c->h_koff,
                                   c->d_twi, L.cand, ntile_n, L.norm, c->d_tref, L.st)) return 1;
        } else if (cra_launch_ccf(L.spec, map.nrows, c->d_refspec, c->R, c->d_tab, c->htab, c->d_twi, L.cand, ntile_n, L.st)) return 1;
        if (tm) CRA_CUDA(cudaEventRecord(c->ev[4 * bi + 2], L.st));
        if (cra_launch_finalize(L.spec, c->d_refspec, c->R, c->d_tab, c->htab, L.cand, ntile_n, map,
                                c->d_res + bfirst[bi], c->fmt, c->frag, c->d_twd, L.norm, c->d_tref, L.st)) return 1;
        if (tm) CRA_CUDA(cudaEventRecord(c->ev[4 * bi + 3], L.st));
        launches += 3;
        c->last_rows = map.nrows; c->last_group = bgroup[bi]; c->last_spec = L.spec; c->last_norm = L.norm;
    }
    if (two_lanes) {                                 // the main stream continues after lane 2 has drained
        CRA_CUDA(cudaEventRecord(c->ev_lane[1], c->st2));
        CRA_CUDA(cudaStreamWaitEvent(c->st, c->ev_lane[1], 0));
    }

    CRA_CUDA(cudaMemcpyAsync(c->h_res, c->d_res, (size_t)n * sizeof(CraResult), cudaMemcpyDeviceToHost, c->st));
    CRA_CUDA(cudaStreamSynchronize(c->st));
    memcpy(out, c->h_res, (size_t)n * sizeof(CraResult));
    if (class_of) for (int p = 0; p < n; ++p) out[p].iref = class_of[p];     // the kernels saw reference 0 of their run
    c->stats = CraAlignStats{};
    c->stats.launches = launches;
    c->stats.rows = total_rows;
    c->stats.alignments = class_of ? total_rows : total_rows * c->R;
    if (tm) {
        for (size_t bi = 0; bi < nb; ++bi) {
            float a = 0, bb = 0, cc = 0;
            cudaEventElapsedTime(&a, c->ev[4 * bi + 0], c->ev[4 * bi + 1]);
            cudaEventElapsedTime(&bb, c->ev[4 * bi + 1], c->ev[4 * bi + 2]);
            cudaEventElapsedTime(&cc, c->ev[4 * bi + 2], c->ev[4 * bi + 3]);
            c->stats.ms_polar += a; c->stats.ms_ccf += bb; c->stats.ms_final += cc;
        }
        float tot = 0; cudaEventElapsedTime(&tot, c->ev[0], c->ev[4 * (nb - 1) + 3]);
        c->stats.ms_total = tot;
    }
    return 0;
}

extern "C" int cra_align(CraCtx* c, int start, int stop, const CraSearch* search, CraResult* out)
{ return align_impl(c, start, stop, search, nullptr, out); }

extern "C" int cra_align_bound(CraCtx* c, int start, int stop, const CraSearch* search, const int* class_of, CraResult* out)
{
    if (!class_of) { cra_set_error("class_of is null"); return 1; }
    return align_impl(c, start, stop, search, class_of, out);
}

extern "C" int cra_last_align_stats(CraCtx* c, CraAlignStats* out)
{
    if (!c || !out) { cra_set_error("null argument"); return 1; }
    *out = c->stats; return 0;
}
extern "C" int cra_set_timing(CraCtx* c, int enabled)
{
    if (!c) { cra_set_error("null context"); return 1; }
    c->timing = enabled != 0; return 0;
}

extern "C" int cra_set_normalize_ring(CraCtx* c, int enabled)
{
    if (!c) { cra_set_error("null context"); return 1; }
    c->cfg.normalize_ring = enabled != 0; return 0;
}
extern "C" int cra_set_step(CraCtx* c, float step)
{
    if (!c) { cra_set_error("null context"); return 1; }
    if (!(step > 0.f)) { cra_set_error("step must be positive"); return 1; }
    c->cfg.step = step; return 0;
}
extern "C" int cra_row_batch(CraCtx* c) { return c ? c->row_batch : 0; }
extern "C" int cra_device_images_ptr(CraCtx* c, void** p) { if (!c || !p) { cra_set_error("null argument"); return 1; } *p = c->d_images; return 0; }
extern "C" void* cra_stream(CraCtx* c) { return c ? (void*)c->st : nullptr; }

extern "C" int cra_zero_sums(CraCtx* c)
{
    Bind b(c); if (b.ok()) return 1;
    const size_t nsum = (size_t)c->cfg.max_refs * 2 * c->npix + c->cfg.max_refs;
    CRA_CUDA(cudaMemsetAsync(c->d_sums, 0, nsum * sizeof(float), c->st));
    CRA_CUDA(cudaStreamSynchronize(c->st));
    return 0;
}

// (alpha [deg], sx, sy, mirror) -> what rotsum_kernel takes: the angle in radians as a float, formed the way EMAN2
// does: EMData::rot_scale_trans2D_background(float angDeg, ...) receives the degrees ALREADY rounded to float (the
// Python double passes through a float parameter) and computes  float ang = angDeg * M_PI / 180.0f  in double.
template <typename T>
static inline float4 pack_par(const T* p)
{
    const float deg = (float)p[0];
    return make_float4((float)((double)deg * 3.14159265358979323846 / 180.0), (float)p[1], (float)p[2], (float)p[3]);
}

template <typename T>
static int accumulate_impl(CraCtx* c, int start, int stop, const T* params, const int* iref, long goff)
{
    CraNvtx range("cra_accumulate (rot_shift2D + class sums)");
    Bind b(c); if (b.ok()) return 1;
    const int n = stop - start;
    if (start < 0 || n < 0 || stop > c->cfg.max_particles) { cra_set_error("particle range exceeds max_particles"); return 1; }
    if (n == 0) return 0;
    if (wait_uploads(c, start, n)) return 1;
    if (ensure_par(c, n)) return 1;
    for (int i = 0; i < n; ++i) {
        if (iref[i] >= c->cfg.max_refs) { cra_set_error("iref exceeds max_refs"); return 1; }
        c->h_par[i] = pack_par(params + 4 * (size_t)i);
        c->h_iref[i] = iref[i];
    }
    CRA_CUDA(cudaMemcpyAsync(c->d_par, c->h_par, sizeof(float4) * n, cudaMemcpyHostToDevice, c->st));
    CRA_CUDA(cudaMemcpyAsync(c->d_iref, c->h_iref, sizeof(int) * n, cudaMemcpyHostToDevice, c->st));
    float* counts = c->d_sums + (size_t)c->cfg.max_refs * 2 * c->npix;
    if (cra_launch_rotsum(c->d_images, c->nx, start, n, c->d_par, c->d_iref, goff, c->d_sums, counts, nullptr, c->st)) return 1;
    CRA_CUDA(cudaStreamSynchronize(c->st));
    return 0;
}

extern "C" int cra_accumulate(CraCtx* c, int start, int stop, const float* params, const int* iref, long goff)
{ return accumulate_impl(c, start, stop, params, iref, goff); }
extern "C" int cra_accumulate_d(CraCtx* c, int start, int stop, const double* params, const int* iref, long goff)
{ return accumulate_impl(c, start, stop, params, iref, goff); }

extern "C" int cra_sums_device_ptr(CraCtx* c, void** dev_ptr, size_t* n_floats)
{
    if (!c) { cra_set_error("null context"); return 1; }
    if (dev_ptr) *dev_ptr = c->d_sums;
    if (n_floats) *n_floats = (size_t)c->cfg.max_refs * 2 * c->npix + c->cfg.max_refs;
    return 0;
}

extern "C" int cra_get_sums(CraCtx* c, float* host_sums, float* host_counts)
{
    Bind b(c); if (b.ok()) return 1;
    const int R = c->cfg.max_refs;
    if (host_sums) CRA_CUDA(cudaMemcpyAsync(host_sums, c->d_sums, (size_t)R * 2 * c->npix * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    if (host_counts) CRA_CUDA(cudaMemcpyAsync(host_counts, c->d_sums + (size_t)R * 2 * c->npix, R * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    CRA_CUDA(cudaStreamSynchronize(c->st));
    return 0;
}

template <typename T>
static int transform_dev_impl(CraCtx* c, int start, int stop, const T* params, float* dev_out)
{
    Bind b(c); if (b.ok()) return 1;
    const int n = stop - start;
    if (start < 0 || n < 0 || stop > c->cfg.max_particles) { cra_set_error("particle range exceeds max_particles"); return 1; }
    if (n == 0) return 0;
    if (!dev_out) { cra_set_error("null output buffer"); return 1; }
    if (wait_uploads(c, start, n)) return 1;
    const int chunk = std::min(n, 65536);
    if (ensure_par(c, chunk)) return 1;
    for (int s = 0; s < n; s += chunk) {
        const int m = std::min(chunk, n - s);
        for (int i = 0; i < m; ++i) c->h_par[i] = pack_par(params + 4 * (size_t)(s + i));
        CRA_CUDA(cudaMemcpyAsync(c->d_par, c->h_par, sizeof(float4) * m, cudaMemcpyHostToDevice, c->st));
        if (cra_launch_rotsum(c->d_images, c->nx, start + s, m, c->d_par, nullptr, 0, nullptr, nullptr,
                              dev_out + (size_t)s * c->npix, c->st)) return 1;
        CRA_CUDA(cudaStreamSynchronize(c->st));       // h_par is reused by the next chunk
    }
    return 0;
}

extern "C" int cra_transform_dev(CraCtx* c, int start, int stop, const float* params, float* dev_out)
{ return transform_dev_impl(c, start, stop, params, dev_out); }

template <typename T>
static int transform_impl(CraCtx* c, int start, int stop, const T* params, float* host_out)
{
    Bind b(c); if (b.ok()) return 1;
    const int n = stop - start;
    if (start < 0 || n < 0 || stop > c->cfg.max_particles) { cra_set_error("particle range exceeds max_particles"); return 1; }
    if (n == 0) return 0;
    if (wait_uploads(c, start, n)) return 1;
    const int chunk = std::min(n, 8192);
    if ((size_t)chunk > c->cap_tmpimg) {
        if (c->d_tmpimg) cudaFree(c->d_tmpimg);
        c->cap_tmpimg = chunk;
        CRA_CUDA(cudaMalloc(&c->d_tmpimg, (size_t)chunk * c->npix * sizeof(float)));
    }
    if (ensure_par(c, chunk)) return 1;
    for (int s = 0; s < n; s += chunk) {
        const int m = std::min(chunk, n - s);
        for (int i = 0; i < m; ++i) c->h_par[i] = pack_par(params + 4 * (size_t)(s + i));
        CRA_CUDA(cudaMemcpyAsync(c->d_par, c->h_par, sizeof(float4) * m, cudaMemcpyHostToDevice, c->st));
        if (cra_launch_rotsum(c->d_images, c->nx, start + s, m, c->d_par, nullptr, 0, nullptr, nullptr, c->d_tmpimg, c->st)) return 1;
        CRA_CUDA(cudaMemcpyAsync(host_out + (size_t)s * c->npix, c->d_tmpimg, (size_t)m * c->npix * sizeof(float), cudaMemcpyDeviceToHost, c->st));
        CRA_CUDA(cudaStreamSynchronize(c->st));
    }
    return 0;
}

extern "C" int cra_transform(CraCtx* c, int start, int stop, const float* params, float* host_out)
{ return transform_impl(c, start, stop, params, host_out); }
extern "C" int cra_transform_d(CraCtx* c, int start, int stop, const double* params, float* host_out)
{ return transform_impl(c, start, stop, params, host_out); }

// device spectrum staged in h_group -> SPIDER packed layout.  F32: row `lane4` of the 4-row group;
// FRAG: the staged row itself (value = bf16 hi + bf16 lo).
static void unpack_spectrum(const CraCtx* c, int lane4, float* out, bool row_layout)
{
    const CraRingTab& t = c->htab;
    const unsigned char* fb = reinterpret_cast<const unsigned char*>(c->h_group);
    for (int i = 0; i < t.nring; ++i) {
        const int half = t.len[i] >> 1;
        float* o = out + t.off[i];
        for (int k = 0; k <= half; ++k) {
            float2 v;
            if (c->fmt == CRA_FMT_FRAG) {
                const int s = t.nring - 1 - i, gc = c->h_koff[k] + (s >> 4), tq = (s & 15) >> 2, j = s & 3;
                const unsigned short* u = reinterpret_cast<const unsigned short*>(fb + (size_t)gc * 128 + tq * 32);
                auto bf = [](unsigned short h) { unsigned int b = (unsigned int)h << 16; float f; memcpy(&f, &b, 4); return f; };
                if (row_layout && !c->frag.unit_rows) {     // hi{re01, im01, re23, im23}, lo{same}
                    const int w = 4 * (j >> 1) + (j & 1);
                    v.x = bf(u[w]) + bf(u[8 + w]);
                    v.y = bf(u[2 + w]) + bf(u[10 + w]);
                } else {              // [re unit | im unit], unit = hi x4, lo x4
                    v.x = bf(u[j]) + bf(u[4 + j]);
                    v.y = bf(u[8 + j]) + bf(u[12 + j]);
                }
            } else v = c->h_group[cra_spec_idx(t.coff[i], half, lane4, k)];
            if (k == 0) o[0] = v.x;
            else if (k == half) o[1] = v.x;
            else { o[2 * k] = v.x; o[2 * k + 1] = v.y; }
        }
    }
}

extern "C" int cra_polar_spectrum(CraCtx* c, int particle, float cx, float cy, float* host_out)
{
    Bind b(c); if (b.ok()) return 1;
    if (particle < 0 || particle >= c->cfg.max_particles) { cra_set_error("bad particle index"); return 1; }
    if (wait_uploads(c, particle, 1)) return 1;
    if (cra_launch_polar_single(c->d_images + (size_t)particle * c->npix, c->nx, c->d_tab, c->htab, c->d_samp, c->d_sampw,
                                c->d_twf, c->items, cx, cy, c->cfg.normalize_ring, c->d_spec, c->fmt, c->frag, c->d_norm, c->gimg_one, c->st)) return 1;
    const size_t bytes = (c->fmt == CRA_FMT_FRAG) ? c->row_bytes : 4 * c->row_bytes;
    CRA_CUDA(cudaMemcpyAsync(c->h_group, c->d_spec, bytes, cudaMemcpyDeviceToHost, c->st));
    CRA_CUDA(cudaStreamSynchronize(c->st));
    unpack_spectrum(c, 0, host_out, true);
    return 0;
}

extern "C" int cra_batch_row_spectrum(CraCtx* c, int row, float* host_out, int* which_kernel)
{
    Bind b(c); if (b.ok()) return 1;
    if (c->fmt != CRA_FMT_FRAG) { cra_set_error("cra_batch_row_spectrum needs the fragment format"); return 1; }
    if (row < 0 || row >= c->last_rows) { cra_set_error("row outside the last batch"); return 1; }
    float2 nm;
    CRA_CUDA(cudaMemcpyAsync(c->h_group, reinterpret_cast<const unsigned char*>(c->last_spec ? c->last_spec : c->d_spec) + (size_t)row * c->row_bytes,
                             c->row_bytes, cudaMemcpyDeviceToHost, c->st));
    CRA_CUDA(cudaMemcpyAsync(&nm, (c->last_norm ? c->last_norm : c->d_norm) + row, sizeof(float2), cudaMemcpyDeviceToHost, c->st));
    CRA_CUDA(cudaStreamSynchronize(c->st));
    unpack_spectrum(c, 0, host_out, true);
    // deferred Normalize_ring: DC bin of every ring -= avg * len, everything * 1/sigma
    const CraRingTab& t = c->htab;
    for (int i = 0; i < t.nring; ++i) host_out[t.off[i]] -= nm.x * (float)t.len[i];
    for (int i = 0; i < t.lcirc; ++i) host_out[i] *= nm.y;
    if (which_kernel) *which_kernel = c->last_group;
    return 0;
}

extern "C" int cra_ref_spectrum(CraCtx* c, int iref, float* host_out)
{
    Bind b(c); if (b.ok()) return 1;
    if (iref < 0 || iref >= c->R) { cra_set_error("bad reference index"); return 1; }
    const unsigned char* src = reinterpret_cast<const unsigned char*>(c->d_refspec);
    if (c->fmt == CRA_FMT_FRAG)
        CRA_CUDA(cudaMemcpyAsync(c->h_group, src + (size_t)iref * c->row_bytes, c->row_bytes, cudaMemcpyDeviceToHost, c->st));
    else
        CRA_CUDA(cudaMemcpyAsync(c->h_group, src + (size_t)(iref >> 2) * 4 * c->row_bytes, 4 * c->row_bytes, cudaMemcpyDeviceToHost, c->st));
    CRA_CUDA(cudaStreamSynchronize(c->st));
    unpack_spectrum(c, iref & 3, host_out, false);
    return 0;
}

extern "C" int cra_ccf_curves(CraCtx* c, int particle, float cx, float cy, int iref, float* q_out, float* t_out)
{
    Bind b(c); if (b.ok()) return 1;
    if (particle < 0 || particle >= c->cfg.max_particles || iref < 0 || iref >= c->R) { cra_set_error("bad index"); return 1; }
    if (wait_uploads(c, particle, 1)) return 1;
    if (cra_launch_polar_single(c->d_images + (size_t)particle * c->npix, c->nx, c->d_tab, c->htab, c->d_samp, c->d_sampw,
                                c->d_twf, c->items, cx, cy, c->cfg.normalize_ring, c->d_spec, c->fmt, c->frag, c->d_norm, c->gimg_one, c->st)) return 1;
    if (cra_launch_ccf_curves(c->d_spec, 0, c->d_refspec, iref, c->d_tab, c->htab,
                              c->d_curves, c->d_curves + c->htab.maxrin, c->fmt, c->frag, c->st)) return 1;
    CRA_CUDA(cudaMemcpyAsync(q_out, c->d_curves, c->htab.maxrin * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    CRA_CUDA(cudaMemcpyAsync(t_out, c->d_curves + c->htab.maxrin, c->htab.maxrin * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    CRA_CUDA(cudaStreamSynchronize(c->st));
    return 0;
}

extern "C" int cra_measure_fp32_peak(int device, double* tf_ffma, double* tf_ffma2)
{
    CRA_CUDA(cudaSetDevice(device));
    double a = 0, b = 0;
    if (cra_fp32_peak(&a, &b)) return 1;
    if (tf_ffma) *tf_ffma = a;
    if (tf_ffma2) *tf_ffma2 = b;
    return 0;
}
