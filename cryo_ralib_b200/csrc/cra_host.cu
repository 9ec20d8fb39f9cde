// cra_host.cu -- the per-particle bookkeeping of the reference's Python loop, batched in native
// code (test_mref.py:184-198 and :206; sp_utilities.combine_params2 / inverse_transform2 on EMAN2's
// float32 2-D Transform, sp_alignment.search_range).  Pure host arithmetic: no CUDA calls, so the
// CPU tests can compare it bit for bit with cryo_ralib_b200/alignment.py, whose Transform algebra is
// pinned on the reference's golden values (cuda/EMAN2_test.ipynb cells 23-25).
// Every float operation is written as its own statement in the order the numpy restatement
// evaluates it (compile with FP contraction off), so the results match to the last bit wherever
// libm's cos/sin/atan2 agree with numpy's.
#include "cra_common.cuh"
#include <math.h>
#include <sched.h>
#include <stdlib.h>

// Threads of the host loops (bookkeeping, pinned gathers).  torchrun exports OMP_NUM_THREADS=1 to every rank, which
// would serialise ~30 ms of per-iteration trigonometry at 100k particles per GPU (the +10 ms of a step at N >= 2 in
// round 1's scaling run): the loops therefore name their own team size -- CRA_HOST_THREADS, else the cores this
// process may run on divided among the ranks of the node (LOCAL_WORLD_SIZE), at most 16.
int cra_host_threads()
{
    static int cached = 0;
    if (cached) return cached;
    int n = 0;
    if (const char* e = getenv("CRA_HOST_THREADS")) n = atoi(e);
    if (n <= 0) {
        cpu_set_t set; CPU_ZERO(&set);
        int cores = (sched_getaffinity(0, sizeof(set), &set) == 0) ? CPU_COUNT(&set) : 1;
        int ranks = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(e) > 0 ? atoi(e) : 1;
        n = cores / ranks;
        if (n > 16) n = 16;
    }
    if (n < 1) n = 1;
    cached = n;
    return n;
}

namespace {

struct M23 { float m[2][3]; };

inline M23 make_t(double alpha_deg, double tx, double ty, int mirror)
{
    const double a = alpha_deg * M_PI / 180.0;
    const float c = (float)cos(a), s = (float)sin(a);
    M23 t;
    t.m[0][0] = c;  t.m[0][1] = s; t.m[0][2] = (float)tx;
    t.m[1][0] = -s; t.m[1][1] = c; t.m[1][2] = (float)ty;
    if (mirror) { t.m[0][0] = -t.m[0][0]; t.m[0][1] = -t.m[0][1]; t.m[0][2] = -t.m[0][2]; }
    return t;
}

inline M23 mul_t(const M23& a, const M23& b)      // a * b, float32 products and sums, left to right
{
    M23 r;
    for (int i = 0; i < 2; ++i) {
        volatile float p, q, u;
        p = a.m[i][0] * b.m[0][0]; q = a.m[i][1] * b.m[1][0]; r.m[i][0] = p + q;
        p = a.m[i][0] * b.m[0][1]; q = a.m[i][1] * b.m[1][1]; r.m[i][1] = p + q;
        p = a.m[i][0] * b.m[0][2]; q = a.m[i][1] * b.m[1][2]; u = p + q; r.m[i][2] = u + a.m[i][2];
    }
    return r;
}

inline void params_t(M23 t, double* alpha, double* sx, double* sy, int* mirror)
{
    volatile float p = t.m[0][0] * t.m[1][1], q = t.m[0][1] * t.m[1][0];
    const float det = p - q;
    const int mir = det < 0.0f;
    if (mir) { t.m[0][0] = -t.m[0][0]; t.m[0][1] = -t.m[0][1]; t.m[0][2] = -t.m[0][2]; }
    double a = atan2((double)t.m[0][1], (double)t.m[0][0]) * (180.0 / M_PI);
    if (a < 0.0) a += 360.0;
    if (a >= 360.0) a -= 360.0;
    *alpha = a; *sx = (double)t.m[0][2]; *sy = (double)t.m[1][2]; *mirror = mir;
}

inline M23 invert_t(const M23& t)                  // Transform::invert: double arithmetic on the float entries
{
    const double m00 = t.m[0][0], m01 = t.m[0][1], m02 = t.m[0][2], m10 = t.m[1][0], m11 = t.m[1][1], m12 = t.m[1][2];
    const double det = m00 * m11 - m01 * m10;
    const double r00 = m11 / det, r01 = -m01 / det, r10 = -m10 / det, r11 = m00 / det;
    const double r02 = -(r00 * m02 + r01 * m12), r12 = -(r10 * m02 + r11 * m12);
    M23 r;
    r.m[0][0] = (float)r00; r.m[0][1] = (float)r01; r.m[0][2] = (float)r02;
    r.m[1][0] = (float)r10; r.m[1][1] = (float)r11; r.m[1][2] = (float)r12;
    return r;
}

inline void search_range(int n, int radius, double shift, double rng, float* left, float* right)
{
    const int cn = n / 2 + 1;
    double ql = cn + shift - radius - 2; if (ql < 0.0) ql = 0.0;
    double qe = n - cn - shift - radius; if (qe < 0.0) qe = 0.0;
    *left = (float)(ql < rng ? ql : rng); *right = (float)(qe < rng ? qe : rng);
}

}  // namespace

// test_mref.py:184-198 for n particles.  params [n][4] double (alpha, sx, sy, mirror) is reset to zero
// in place where the inverted shift exceeds mashi = cnx - ou - 2.
extern "C" int cra_mref_search_request(int n, double* params, int nx, int ou, double xr, double yr,
                                       CraSearch* search, double* sxi_out, double* syi_out)
{
    if (n < 0 || !params || !search || !sxi_out || !syi_out) { cra_set_error("null argument"); return 1; }
    const int cnx = nx / 2 + 1, mashi = cnx - ou - 2;
#pragma omp parallel for schedule(static) num_threads(cra_host_threads()) if (n > 4096)
    for (int i = 0; i < n; ++i) {
        double a, sxi, syi; int mir;
        params_t(invert_t(make_t(params[4 * i], params[4 * i + 1], params[4 * i + 2], 0)), &a, &sxi, &syi, &mir);
        if (fabs(sxi) > mashi || fabs(syi) > mashi) {
            sxi = 0.0; syi = 0.0;
            params[4 * i] = params[4 * i + 1] = params[4 * i + 2] = params[4 * i + 3] = 0.0;
        }
        CraSearch s;
        search_range(nx, ou, sxi, xr, &s.xl, &s.xr);
        search_range(nx, ou, syi, yr, &s.yl, &s.yr);
        s.cx = (float)(cnx + sxi); s.cy = (float)(cnx + syi);
        search[i] = s; sxi_out[i] = sxi; syi_out[i] = syi;
    }
    return 0;
}

// ali2d_single_iter's prologue for n particles (test_reffree.py:780-783 -> Sphire): fold the average's centre
// shift cs into the parameters (combine_params2(p, 0, -cs)), invert, clamp the shift to +-mashi.
extern "C" int cra_reffree_search_request(int n, const double* params, double csx, double csy, int nx, int ou,
                                          double xr, double yr, CraSearch* search, double* sxi_out, double* syi_out)
{
    if (n < 0 || !params || !search || !sxi_out || !syi_out) { cra_set_error("null argument"); return 1; }
    const int cnx = nx / 2 + 1, mashi = cnx - ou - 2;
    const M23 t2 = make_t(0.0, -csx, -csy, 0);
#pragma omp parallel for schedule(static) num_threads(cra_host_threads()) if (n > 4096)
    for (int i = 0; i < n; ++i) {
        double a, sx, sy, a2, sxi, syi; int mir, mir2;
        params_t(mul_t(t2, make_t(params[4 * i], params[4 * i + 1], params[4 * i + 2], (int)params[4 * i + 3])), &a, &sx, &sy, &mir);
        params_t(invert_t(make_t(a, sx, sy, 0)), &a2, &sxi, &syi, &mir2);
        if (sxi < -mashi) sxi = -mashi; if (sxi > mashi) sxi = mashi;
        if (syi < -mashi) syi = -mashi; if (syi > mashi) syi = mashi;
        CraSearch s;
        search_range(nx, ou, sxi, xr, &s.xl, &s.xr);
        search_range(nx, ou, syi, yr, &s.yl, &s.yr);
        s.cx = (float)(cnx + sxi); s.cy = (float)(cnx + syi);
        search[i] = s; sxi_out[i] = sxi; syi_out[i] = syi;
    }
    return 0;
}

// test_mref.py:206: combine_params2(0, -sxi, -syi, 0, ang, sxs, sys, mirror) -> params_out [n][4] double
extern "C" int cra_compose_result(int n, const double* sxi, const double* syi, const CraResult* res, double* params_out)
{
    if (n < 0 || !sxi || !syi || !res || !params_out) { cra_set_error("null argument"); return 1; }
#pragma omp parallel for schedule(static) num_threads(cra_host_threads()) if (n > 4096)
    for (int i = 0; i < n; ++i) {
        const M23 t1 = make_t(0.0, -sxi[i], -syi[i], 0);
        const M23 t2 = make_t((double)res[i].ang, (double)res[i].sxs, (double)res[i].sys, res[i].mirror);
        double a, sx, sy; int mir;
        params_t(mul_t(t2, t1), &a, &sx, &sy, &mir);
        params_out[4 * i] = a; params_out[4 * i + 1] = sx; params_out[4 * i + 2] = sy; params_out[4 * i + 3] = (double)mir;
    }
    return 0;
}

// ---- fit_tanh: the Nelder-Mead fit of the tangent low-pass to 2 FSC / (1 + FSC) (sp_filter.fit_tanh ->
// sp_utilities.amoeba; the user function ref_ali2d of test_mref.py:273-276 and test_reffree.py:733 calls it for
// every class and iteration).  Step for step the simplex of cryo_ralib_b200/refupdate.py: amoeba, in double;
// ~1500 evaluations of a 46-term cost, which the interpreted loop needs 30 ms for -- a third of a reference-free
// iteration on this engine.
namespace {
struct TanhFit { int n; const double* freq; const double* target; };
double tanh_cost(const TanhFit& d, double fl, double aa)
{
    if (fl == 0.0 || aa == 0.0) return -0.0;
    const double c = M_PI / 2.0 / aa / fl;
    double acc = 0.0;
    for (int i = 0; i < d.n; ++i) {
        const double qt = d.target[i] - 0.5 * (tanh(c * (d.freq[i] + fl)) - tanh(c * (d.freq[i] - fl)));
        acc += qt * qt;
    }
    return -acc;
}
}  // namespace

extern "C" int cra_fit_tanh(int n, const double* freq, const double* target, double fl0, double aa0,
                            double scale_fl, double scale_aa, double* out4)
{
    if (n < 1 || !freq || !target || !out4) { cra_set_error("cra_fit_tanh: bad arguments"); return 1; }
    const TanhFit d{n, freq, target};
    const double ftol = 1.e-4, xtol = 1.e-4; const int itmax = 500;
    const double scale[2] = {scale_fl, scale_aa};
    double sx[3][2] = {{fl0, aa0}, {fl0 + scale_fl, aa0}, {fl0, aa0 + scale_aa}};
    double fv[3];
    for (int i = 0; i < 3; ++i) fv[i] = tanh_cost(d, sx[i][0], sx[i][1]);
    int iteration = 0;
    for (;;) {
        int worst = 0, best = 0;
        for (int i = 0; i < 3; ++i) { if (fv[i] > fv[best]) best = i; if (fv[i] < fv[worst]) worst = i; }
        double pavg[2];
        for (int j = 0; j < 2; ++j) { double a = 0.0; for (int i = 0; i < 3; ++i) if (i != worst) a += sx[i][j]; pavg[j] = a / 2.0; }
        double simscale = 0.0;
        for (int j = 0; j < 2; ++j) simscale += fabs(pavg[j] - sx[worst][j]) / scale[j];
        simscale /= 2.0;
        const double fscale = (fabs(fv[best]) + fabs(fv[worst])) / 2.0;
        const double frange = fscale != 0.0 ? fabs(fv[best] - fv[worst]) / fscale : 0.0;
        if ((frange < ftol && simscale < xtol) || iteration >= itmax) {
            out4[0] = sx[best][0]; out4[1] = sx[best][1]; out4[2] = fv[best]; out4[3] = (double)iteration;
            return 0;
        }
        double pnew[2] = {2.0 * pavg[0] - sx[worst][0], 2.0 * pavg[1] - sx[worst][1]};
        double fnew = tanh_cost(d, pnew[0], pnew[1]);
        if (fnew <= fv[worst]) {
            for (int i = 0; i < 3; ++i)
                if (i != best && i != worst) {
                    for (int j = 0; j < 2; ++j) sx[i][j] = 0.5 * sx[best][j] + 0.5 * sx[i][j];
                    fv[i] = tanh_cost(d, sx[i][0], sx[i][1]);
                }
            for (int j = 0; j < 2; ++j) pnew[j] = 0.5 * sx[best][j] + 0.5 * sx[worst][j];
            fnew = tanh_cost(d, pnew[0], pnew[1]);
        } else if (fnew >= fv[best]) {
            const double p2[2] = {3.0 * pavg[0] - 2.0 * sx[worst][0], 3.0 * pavg[1] - 2.0 * sx[worst][1]};
            const double f2 = tanh_cost(d, p2[0], p2[1]);
            if (f2 > fnew) { pnew[0] = p2[0]; pnew[1] = p2[1]; fnew = f2; }
        }
        sx[worst][0] = pnew[0]; sx[worst][1] = pnew[1];
        fv[worst] = fnew;
        ++iteration;
    }
}
