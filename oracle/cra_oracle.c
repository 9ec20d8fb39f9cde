/*
 * cra_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * A plain-C restatement of the EMAN2 2.31 / Sphire CPU algorithm that the
 * reference's CPU drivers call for 2D multi-reference / reference-free
 * alignment (reference call sites: test_mref.py:145-146, :171-175, :184-215;
 * test_reffree.py:780-783).  The arithmetic itself lives in the un-vendored
 * third-party dependency EMAN2 2.31 (README.md:43: libEM/sparx/util_sparx.cpp,
 * libEM/emdata_sparx.cpp, libEM/processor.cpp, libEM/transform.cpp and Sphire's
 * sp_alignment.py / sp_utilities.py / sp_fundamentals.py); none of it is under
 * /root/reference, so this file restates the published algorithm (SURVEY.md
 * Appendix A) and is cross-checked against the restatements the reference does
 * carry: quadri_background / rot_scale_trans2D_background / mirror
 * (notebook/02_CuPy_Image_Processing_rot_shift2d.ipynb cell 2), the prb1d
 * coefficients (cuda/gpu_aln_noref.cu:1438-1439), the final shift rotation
 * (test_mref_gpu_align.py:578-588) and the Transform known answers
 * (cuda/EMAN2_test.ipynb cells 23-25).
 *
 * PARITY STATUS: the Transform algebra is pinned by the reference's three golden
 * tuples.  At the multiref_polar_ali_2d boundary the reference holds no EMAN2
 * vectors and EMAN2 cannot be run here: the DIGITS of this restatement are
 * "parity unpinned" and carried by the self-consistency tests of
 * tests/test_oracle.py.  Its CONVENTIONS and discrete answers (class, mirror flag,
 * integer shift, angle direction and origin) are pinned to output of the
 * reference itself: AlignParam[] as the reference's own CUDA library, compiled
 * unchanged for sm_100, left it for a deterministic stack
 * (tests/golden/refcuda_mref_outputs.npz, tests/golden/make_refcuda_case.py):
 * 512 / 512 particles identical, angles within 0.12 degrees.  That library is
 * gpu_isac's arithmetic, not EMAN2's, so this is not a last-digit pin.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (cryo_ralib_b200)
 * never does.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CRA_PI 3.14159265358979323846

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ */
/* A.1 Numrinit / ringwe  (sp_alignment.py; called test_mref.py:145-146) */

static int ilog2_floor(int n) { int l = -1; while (n > 0) { n >>= 1; ++l; } return l; }

/* numr holds (radius, 1-based offset, length) triplets; returns nring. */
int cra_o_numrinit(int first_ring, int last_ring, int skip, int *numr, int cap_rings)
{
    const int MAXFFT = 32768;
    int lcirc = 1, n = 0;
    for (int k = first_ring; k <= last_ring; k += skip) {
        if (n >= cap_rings) return -1;
        int jp = (int)(2.0 * CRA_PI * k + 0.5);
        int ip = 1 << (ilog2_floor(jp) + 1);
        if (k + skip <= last_ring && jp > ip + ip / 2) ip = (2 * ip < MAXFFT) ? 2 * ip : MAXFFT;
        if (k + skip > last_ring && jp > ip + ip / 5) ip = (2 * ip < MAXFFT) ? 2 * ip : MAXFFT;
        numr[3 * n] = k; numr[3 * n + 1] = lcirc; numr[3 * n + 2] = ip;
        lcirc += ip; ++n;
    }
    return n;
}

void cra_o_ringwe(const int *numr, int nring, float *wr)
{
    double maxrin = (double)numr[3 * nring - 1];
    for (int i = 0; i < nring; ++i) {
        double L = (double)numr[3 * i + 2];
        wr[i] = (float)(numr[3 * i] * (2.0 * CRA_PI) / L * maxrin / L);
    }
}

/* ------------------------------------------------------------------ */
/* A.2 quadri / Polar2Dm (util_sparx.cpp Util::quadri, Util::alrl_ms)  */

#define FD(i, j) fdata[((i) - 1) + (size_t)((j) - 1) * nxdata]

static float quadri(float xx, float yy, int nxdata, int nydata, const float *fdata)
{
    float x = xx, y = yy;
    while (x < 1.0f) x += nxdata;
    while (x >= (float)(nxdata + 1)) x -= nxdata;
    while (y < 1.0f) y += nydata;
    while (y >= (float)(nydata + 1)) y -= nydata;
    int i = (int)x, j = (int)y;
    float dx0 = x - i, dy0 = y - j;
    int ip1 = i + 1, im1 = i - 1, jp1 = j + 1, jm1 = j - 1;
    if (ip1 > nxdata) ip1 -= nxdata;
    if (im1 < 1) im1 += nxdata;
    if (jp1 > nydata) jp1 -= nydata;
    if (jm1 < 1) jm1 += nydata;
    float f0 = FD(i, j);
    float c1 = FD(ip1, j) - f0;
    float c2 = (c1 - f0 + FD(im1, j)) * 0.5f;
    float c3 = FD(i, jp1) - f0;
    float c4 = (c3 - f0 + FD(i, jm1)) * 0.5f;
    float dxb = dx0 - 1, dyb = dy0 - 1;
    int hxc = (dx0 >= 0) ? 1 : -1, hyc = (dy0 >= 0) ? 1 : -1;
    int ic = i + hxc, jc = j + hyc;
    if (ic > nxdata) ic -= nxdata; else if (ic < 1) ic += nxdata;
    if (jc > nydata) jc -= nydata; else if (jc < 1) jc += nydata;
    float c5 = ((FD(ic, jc) - f0 - hxc * c1 - (hxc * (hxc - 1.0f)) * c2
                 - hyc * c3 - (hyc * (hyc - 1.0f)) * c4) * (hxc * hyc));
    return f0 + dx0 * (c1 + dxb * c2 + dy0 * c5) + dy0 * (c3 + dyb * c4);
}

/* quadri_background: out-of-frame sample point replaced by the output pixel's
 * own position (cross-check: notebook/02 cell 2, quadri_background). */
static float quadri_background(float xx, float yy, int nxdata, int nydata,
                               const float *fdata, int xnew, int ynew)
{
    float x = xx, y = yy;
    if ((x < 1.0f) || (x >= (float)(nxdata + 1)) || (y < 1.0f) || (y >= (float)(nydata + 1))) {
        x = (float)xnew; y = (float)ynew;
    }
    int i = (int)x, j = (int)y;
    float dx0 = x - i, dy0 = y - j;
    int ip1 = i + 1, im1 = i - 1, jp1 = j + 1, jm1 = j - 1;
    if (ip1 > nxdata) ip1 -= nxdata;
    if (im1 < 1) im1 += nxdata;
    if (jp1 > nydata) jp1 -= nydata;
    if (jm1 < 1) jm1 += nydata;
    float f0 = FD(i, j);
    float c1 = FD(ip1, j) - f0;
    float c2 = (c1 - f0 + FD(im1, j)) * 0.5f;
    float c3 = FD(i, jp1) - f0;
    float c4 = (c3 - f0 + FD(i, jm1)) * 0.5f;
    float dxb = dx0 - 1, dyb = dy0 - 1;
    int hxc = (dx0 >= 0) ? 1 : -1, hyc = (dy0 >= 0) ? 1 : -1;
    int ic = i + hxc, jc = j + hyc;
    if (ic > nxdata) ic -= nxdata; else if (ic < 1) ic += nxdata;
    if (jc > nydata) jc -= nydata; else if (jc < 1) jc += nydata;
    float c5 = ((FD(ic, jc) - f0 - hxc * c1 - (hxc * (hxc - 1.0f)) * c2
                 - hyc * c3 - (hyc * (hyc - 1.0f)) * c4) * (hxc * hyc));
    return f0 + dx0 * (c1 + dxb * c2 + dy0 * c5) + dy0 * (c3 + dyb * c4);
}
#undef FD

/* Polar2Dm(image, cns2, cnr2, numr, "F"): circ has lcirc floats. */
void cra_o_polar2dm(const float *xim, int nsam, int nrow, float cns2, float cnr2,
                    const int *numr, int nring, float *circ)
{
    const double dpi = 2 * atan(1.0);
    for (int it = 0; it < nring; ++it) {
        int inr = numr[3 * it];
        int l = numr[3 * it + 2];
        int lt = l / 4;
        int nsim = lt - 1;
        double dfi = dpi / (nsim + 1);
        float *c = circ + (numr[3 * it + 1] - 1);
        c[0]      = quadri(0.0f + cns2, inr + cnr2, nsam, nrow, xim);
        c[lt]     = quadri(inr + cns2, 0.0f + cnr2, nsam, nrow, xim);
        c[2 * lt] = quadri(0.0f + cns2, -inr + cnr2, nsam, nrow, xim);
        c[3 * lt] = quadri(-inr + cns2, 0.0f + cnr2, nsam, nrow, xim);
        for (int jt = 1; jt <= nsim; ++jt) {
            float fi = (float)(dfi * jt);
            float x = sinf(fi) * inr;
            float y = cosf(fi) * inr;
            c[jt]          = quadri(x + cns2, y + cnr2, nsam, nrow, xim);
            c[jt + lt]     = quadri(y + cns2, -x + cnr2, nsam, nrow, xim);
            c[jt + 2 * lt] = quadri(-x + cns2, -y + cnr2, nsam, nrow, xim);
            c[jt + 3 * lt] = quadri(-y + cns2, x + cnr2, nsam, nrow, xim);
        }
    }
}

/* ------------------------------------------------------------------ */
/* A.3 Normalize_ring(circ, numr, 0) (util_sparx.cpp)                  */

void cra_o_normalize_ring(float *circ, const int *numr, int nring)
{
    float av = 0.0f, sq = 0.0f, nn = 0.0f;
    for (int i = 0; i < nring; ++i) {
        int len = numr[3 * i + 2], off = numr[3 * i + 1] - 1;
        float w = (float)(numr[3 * i] * 2 * CRA_PI / (float)len);
        for (int j = 0; j < len; ++j) {
            float v = circ[off + j];
            av += v * w; sq += v * v * w; nn += w;
        }
    }
    float avg = av / nn;
    float sgm = sqrtf((sq - av * av / nn) / nn);
    int lcirc = numr[3 * nring - 2] + numr[3 * nring - 1] - 1;
    for (int i = 0; i < lcirc; ++i) { circ[i] -= avg; circ[i] /= sgm; }
}

/* ------------------------------------------------------------------ */
/* A.4 Frngs: per-ring real FFT in SPIDER packed layout.
 * slot0 = Re F_0, slot1 = Re F_{L/2}, slots (2k,2k+1) = (Re,Im) F_k, with
 * F_k = sum_n x_n exp(-2 pi i n k / L) (unnormalised forward; the inverse
 * divides by L).  The correlation produced from it is independent of the sign
 * of the exponent as long as forward and inverse agree (SURVEY A.9 ix).
 * Arithmetic class follows fftr_q (float) / fftr_d (double).             */

/* twiddle tables, one per power-of-two length (built once; cos/sin of 2 pi k / n, k < n/2) */
#define CRA_MAXLOG 14
static double *g_twc[CRA_MAXLOG + 1], *g_tws[CRA_MAXLOG + 1];
static volatile int g_tw_ready = 0;
static void ensure_tables(void)
{
    if (g_tw_ready) return;
#ifdef _OPENMP
#pragma omp critical(cra_tw)
#endif
    {
        if (!g_tw_ready) {
            for (int lg = 1; lg <= CRA_MAXLOG; ++lg) {
                int n = 1 << lg;
                g_twc[lg] = (double *)malloc(sizeof(double) * (n / 2));
                g_tws[lg] = (double *)malloc(sizeof(double) * (n / 2));
                for (int k = 0; k < n / 2; ++k) {
                    g_twc[lg][k] = cos(2.0 * CRA_PI * k / n);
                    g_tws[lg][k] = sin(2.0 * CRA_PI * k / n);
                }
            }
            g_tw_ready = 1;
        }
    }
}

#define DEF_CFFT(NAME, T)                                                        \
static void NAME(T *re, T *im, int n, int sign)                                  \
{                                                                                \
    for (int i = 1, j = 0; i < n; ++i) {                                         \
        int bit = n >> 1;                                                        \
        for (; j & bit; bit >>= 1) j ^= bit;                                     \
        j ^= bit;                                                                \
        if (i < j) { T t = re[i]; re[i] = re[j]; re[j] = t;                      \
                     t = im[i]; im[i] = im[j]; im[j] = t; }                      \
    }                                                                            \
    int lg = 1;                                                                  \
    for (int len = 2; len <= n; len <<= 1, ++lg) {                               \
        int half = len >> 1;                                                     \
        const double *tc = g_twc[lg], *ts = g_tws[lg];                           \
        for (int k = 0; k < half; ++k) {                                         \
            T wr = (T)tc[k], wi = (T)(sign * ts[k]);                             \
            for (int s = k; s < n; s += len) {                                   \
                int e = s + half;                                                \
                T tr = re[e] * wr - im[e] * wi;                                  \
                T ti = re[e] * wi + im[e] * wr;                                  \
                re[e] = re[s] - tr; im[e] = im[s] - ti;                          \
                re[s] += tr; im[s] += ti;                                        \
            }                                                                    \
        }                                                                        \
    }                                                                            \
}
DEF_CFFT(cfft_f, float)
DEF_CFFT(cfft_d, double)

static int ilog2_exact(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

/* forward real FFT of x[0..L-1] (float), in place, packed layout */
static void rfft_packed_f(float *x, int L)
{
    ensure_tables();
    int n = L / 2;
    float re[4096], im[4096];   /* L <= 8192 */
    for (int i = 0; i < n; ++i) { re[i] = x[2 * i]; im[i] = x[2 * i + 1]; }
    cfft_f(re, im, n, -1);
    /* split: F_k = (Z_k + conj(Z_{n-k}))/2 + e^{-2 pi i k/L} (Z_k - conj(Z_{n-k}))/(2i) */
    x[0] = re[0] + im[0];
    x[1] = re[0] - im[0];
    const double *tc = g_twc[ilog2_exact(L)], *ts = g_tws[ilog2_exact(L)];
    for (int k = 1; k < n; ++k) {
        int m = n - k;
        float er = 0.5f * (re[k] + re[m]), ei = 0.5f * (im[k] - im[m]);
        float orr = 0.5f * (im[k] + im[m]), oi = -0.5f * (re[k] - re[m]);
        float wr = (float)tc[k], wi = (float)(-ts[k]);
        x[2 * k]     = er + (orr * wr - oi * wi);
        x[2 * k + 1] = ei + (orr * wi + oi * wr);
    }
}

/* inverse of the above in double: packed spectrum -> real sequence, 1/L scaled */
static void irfft_packed_d(double *x, int L)
{
    ensure_tables();
    int n = L / 2;
    double re[4096], im[4096]; /* L <= 8192 */
    /* Z_k = E_k + i O_k with E_k = (F_k + conj(F_{n-k}))/2, O_k = e^{+2 pi i k/L}(F_k - conj(F_{n-k}))/2 */
    double f0 = x[0], fn = x[1];
    re[0] = 0.5 * (f0 + fn); im[0] = 0.5 * (f0 - fn);
    const double *tc = g_twc[ilog2_exact(L)], *ts = g_tws[ilog2_exact(L)];
    for (int k = 1; k < n; ++k) {
        int m = n - k;
        double fkr = x[2 * k], fki = x[2 * k + 1];
        double fmr = x[2 * m], fmi = -x[2 * m + 1];          /* conj(F_{n-k}) */
        double er = 0.5 * (fkr + fmr), ei = 0.5 * (fki + fmi);
        double dr = 0.5 * (fkr - fmr), di = 0.5 * (fki - fmi);
        double wr = tc[k], wi = ts[k];
        double orr = dr * wr - di * wi, oi = dr * wi + di * wr;
        re[k] = er - oi; im[k] = ei + orr;                    /* E + i O */
    }
    cfft_d(re, im, n, +1);
    double s = 1.0 / n;
    for (int i = 0; i < n; ++i) { x[2 * i] = re[i] * s; x[2 * i + 1] = im[i] * s; }
}

void cra_o_frngs(float *circ, const int *numr, int nring)
{
    for (int i = 0; i < nring; ++i)
        rfft_packed_f(circ + (numr[3 * i + 1] - 1), numr[3 * i + 2]);
}

/* Applyws (refs only): Nyquist slot halved for rings shorter than maxrin */
void cra_o_applyws(float *circ, const int *numr, int nring, const float *wr)
{
    int maxrin = numr[3 * nring - 1];
    for (int i = 0; i < nring; ++i) {
        int len = numr[3 * i + 2], off = numr[3 * i + 1] - 1;
        float w = wr[i];
        circ[off] *= w;
        if (len == maxrin) circ[off + 1] *= w;
        else circ[off + 1] = (float)(circ[off + 1] * (0.5 * w));
        for (int j = 2 + off; j < len + off; ++j) circ[j] *= w;
    }
}

/* exported for unit tests of the FFT pair */
void cra_o_rfft_packed(float *x, int L) { rfft_packed_f(x, L); }
void cra_o_irfft_packed(double *x, int L) { irfft_packed_d(x, L); }

/* ------------------------------------------------------------------ */
/* A.5 Crosrng_ms + prb1d + ang_n (util_sparx.cpp)                     */

static void prb1d(const double *b, float *pos)
{
    double c2 = 49. * b[0] + 6. * b[1] - 21. * b[2] - 32. * b[3] - 27. * b[4] - 6. * b[5] + 31. * b[6];
    double c3 = 5. * b[0] - 3. * b[2] - 4. * b[3] - 3. * b[4] + 5. * b[6];
    *pos = 0.0f;
    if (c3 != 0.0) *pos = (float)((c2 / (2.0 * c3)) - 4);
}

static float ang_n(float peakp, int maxrin)
{
    return fmodf(((peakp - 1.0f) / maxrin + 1.0f) * 360.0f, 360.0f);
}

/* circ1 = weighted reference spectrum, circ2 = image spectrum.
 * out: qn, tot, qm, tmt; optionally the full q/t curves (maxrin doubles). */
void cra_o_crosrng_ms(const float *circ1, const float *circ2, const int *numr, int nring,
                      double *qn_o, float *tot_o, double *qm_o, float *tmt_o,
                      double *q_curve, double *t_curve)
{
    int maxrin = numr[3 * nring - 1];
    double q[8192 + 2], t[8192 + 2];
    memset(q, 0, sizeof(double) * ((size_t)maxrin + 2));
    memset(t, 0, sizeof(double) * ((size_t)maxrin + 2));
    for (int i = 0; i < nring; ++i) {
        int len = numr[3 * i + 2], off = numr[3 * i + 1] - 1;
        float t1 = circ1[off] * circ2[off];
        q[0] += t1; t[0] += t1;
        t1 = circ1[off + 1] * circ2[off + 1];
        if (len == maxrin) { q[1] += t1; t[1] += t1; }
        else { q[len] += t1; t[len] += t1; }
        for (int j = 2; j < len; j += 2) {
            float c1 = circ1[off + j], c2 = circ1[off + j + 1];
            float d1 = circ2[off + j], d2 = circ2[off + j + 1];
            float p1 = c1 * d1, p2 = c2 * d2, p3 = c1 * d2, p4 = c2 * d1;
            q[j] += p1 + p2;  q[j + 1] += -p3 + p4;
            t[j] += p1 - p2;  t[j + 1] += -p3 - p4;
        }
    }
    irfft_packed_d(q, maxrin);
    irfft_packed_d(t, maxrin);
    double t7[7]; float pos; int jtot = 0;
    double qn = -1.0e20;
    for (int j = 1; j <= maxrin; ++j) if (q[j - 1] >= qn) { qn = q[j - 1]; jtot = j; }
    for (int k = -3; k <= 3; ++k) { int j = ((jtot + k + maxrin - 1) % maxrin) + 1; t7[k + 3] = q[j - 1]; }
    prb1d(t7, &pos);
    *qn_o = qn; *tot_o = (float)jtot + pos;
    double qm = -1.0e20;
    for (int j = 1; j <= maxrin; ++j) if (t[j - 1] >= qm) { qm = t[j - 1]; jtot = j; }
    for (int k = -3; k <= 3; ++k) { int j = ((jtot + k + maxrin - 1) % maxrin) + 1; t7[k + 3] = t[j - 1]; }
    prb1d(t7, &pos);
    *qm_o = qm; *tmt_o = (float)jtot + pos;
    if (q_curve) memcpy(q_curve, q, sizeof(double) * maxrin);
    if (t_curve) memcpy(t_curve, t, sizeof(double) * maxrin);
}

/* ------------------------------------------------------------------ */
/* A.6 multiref_polar_ali_2d (util_sparx.cpp; called test_mref.py:200-201)
 * crefim: R weighted reference spectra, each lcirc floats.
 * out[0..5] = ang, sxs, sys, mirror, nref, peak ; out[6],out[7] = raw -ix,-iy
 * normalize != 0 applies Normalize_ring (multiref path); 0 gives ormq
 * semantics (test_reffree.py:780-783 -> ali2d_single_iter -> ormq).      */
void cra_o_multiref_polar_ali_2d(const float *image, int nx, int ny,
                                 const float *crefim, int R,
                                 float xl, float xr_, float yl, float yr_, float step,
                                 const int *numr, int nring, float cnx, float cny,
                                 int normalize, float *out)
{
    int lcirc = numr[3 * nring - 2] + numr[3 * nring - 1] - 1;
    int maxrin = numr[3 * nring - 1];
    int lkx = (int)(xl / step), rkx = (int)(xr_ / step);
    int lky = (int)(yl / step), rky = (int)(yr_ / step);
    float *cimage = (float *)malloc(sizeof(float) * lcirc);
    float peak = -1.0E23f, ang = 0.0f, sx = 0, sy = 0;
    int nref = 0, mirror = 0;
    for (int i = -lky; i <= rky; ++i) {
        float iy = i * step;
        for (int j = -lkx; j <= rkx; ++j) {
            float ix = j * step;
            cra_o_polar2dm(image, nx, ny, cnx + ix, cny + iy, numr, nring, cimage);
            if (normalize) cra_o_normalize_ring(cimage, numr, nring);
            cra_o_frngs(cimage, numr, nring);
            for (int iref = 0; iref < R; ++iref) {
                double qn, qm; float tot, tmt;
                cra_o_crosrng_ms(crefim + (size_t)iref * lcirc, cimage, numr, nring,
                                 &qn, &tot, &qm, &tmt, NULL, NULL);
                if (qn >= peak || qm >= peak) {
                    sx = -ix; sy = -iy; nref = iref;
                    if (qn >= qm) { ang = ang_n(tot, maxrin); peak = (float)qn; mirror = 0; }
                    else          { ang = ang_n(tmt, maxrin); peak = (float)qm; mirror = 1; }
                }
            }
        }
    }
    free(cimage);
    float co = (float)cos(ang * CRA_PI / 180.0);
    float so = (float)(-sin(ang * CRA_PI / 180.0));
    out[0] = ang; out[1] = sx * co - sy * so; out[2] = sx * so + sy * co;
    out[3] = (float)mirror; out[4] = (float)nref; out[5] = peak;
    out[6] = sx; out[7] = sy;
}

/* ------------------------------------------------------------------ */
/* A.7 2-D Transform algebra (libEM/transform.cpp: float 3x4 matrix,
 * x' = M T R x, get_trans negates tx when mirrored).  Known answers:
 * cuda/EMAN2_test.ipynb cells 23-25.                                  */

typedef struct { float m[2][3]; } T2;

static T2 t2_make(double alpha, double tx, double ty, int mirror)
{
    T2 t; double a = alpha * CRA_PI / 180.0;
    float c = (float)cos(a), s = (float)sin(a);
    t.m[0][0] = c;  t.m[0][1] = s; t.m[0][2] = (float)tx;
    t.m[1][0] = -s; t.m[1][1] = c; t.m[1][2] = (float)ty;
    if (mirror) { t.m[0][0] = -t.m[0][0]; t.m[0][1] = -t.m[0][1]; t.m[0][2] = -t.m[0][2]; }
    return t;
}
static T2 t2_mul(T2 a, T2 b)   /* a*b : apply b first */
{
    T2 r;
    for (int i = 0; i < 2; ++i) {
        r.m[i][0] = a.m[i][0] * b.m[0][0] + a.m[i][1] * b.m[1][0];
        r.m[i][1] = a.m[i][0] * b.m[0][1] + a.m[i][1] * b.m[1][1];
        r.m[i][2] = a.m[i][0] * b.m[0][2] + a.m[i][1] * b.m[1][2] + a.m[i][2];
    }
    return r;
}
static void t2_params(T2 t, double *alpha, double *tx, double *ty, int *mirror)
{
    float det = t.m[0][0] * t.m[1][1] - t.m[0][1] * t.m[1][0];
    int mir = det < 0;
    if (mir) { t.m[0][0] = -t.m[0][0]; t.m[0][1] = -t.m[0][1]; t.m[0][2] = -t.m[0][2]; }
    double a = atan2((double)t.m[0][1], (double)t.m[0][0]) * (180.0 / CRA_PI);   /* EMConsts::rad2deg */
    if (a < 0) a += 360.0;
    if (a >= 360.0) a -= 360.0;
    *alpha = a; *tx = t.m[0][2]; *ty = t.m[1][2]; *mirror = mir;
}
static T2 t2_inverse(T2 t)
{
    /* Transform::invert works in double on the float entries and casts back;
     * this reproduces cuda/EMAN2_test.ipynb cell 25 bit-for-bit. */
    double m00 = t.m[0][0], m01 = t.m[0][1], m10 = t.m[1][0], m11 = t.m[1][1];
    double v0 = t.m[0][2], v1 = t.m[1][2];
    double det = m00 * m11 - m01 * m10;
    double r00 = m11 / det, r01 = -m01 / det, r10 = -m10 / det, r11 = m00 / det;
    T2 r;
    r.m[0][0] = (float)r00; r.m[0][1] = (float)r01;
    r.m[1][0] = (float)r10; r.m[1][1] = (float)r11;
    r.m[0][2] = (float)(-(r00 * v0 + r01 * v1));
    r.m[1][2] = (float)(-(r10 * v0 + r11 * v1));
    return r;
}

void cra_o_combine_params2(double a1, double sx1, double sy1, int m1,
                           double a2, double sx2, double sy2, int m2, double *out4)
{
    T2 t = t2_mul(t2_make(a2, sx2, sy2, m2), t2_make(a1, sx1, sy1, m1));
    int mir; t2_params(t, &out4[0], &out4[1], &out4[2], &mir); out4[3] = mir;
}
void cra_o_inverse_transform2(double alpha, double tx, double ty, int mirror, double *out4)
{
    T2 t = t2_inverse(t2_make(alpha, tx, ty, mirror));
    int mir; t2_params(t, &out4[0], &out4[1], &out4[2], &mir); out4[3] = mir;
}

/* ------------------------------------------------------------------ */
/* A.8 rot_shift2D = rot_scale_trans2D_background (+ xform.mirror x)
 * (emdata_sparx.cpp; cross-check notebook/02 cell 2; called test_mref.py:210) */

static float restrict2(float x, int nx)
{
    while (x >= (float)nx) x -= nx;
    while (x <= -(float)nx) x += nx;
    return x;
}

void cra_o_rot_shift2d(const float *src, int nx, int ny, float angDeg, float delx, float dely,
                       int mirror, float *dst)
{
    float ang = (float)(angDeg * CRA_PI / 180.0);  /* float ang=angDeg*M_PI/180.0f */
    delx = restrict2(delx, nx); dely = restrict2(dely, ny);
    int xc = nx / 2, yc = ny / 2;
    float shiftxc = xc + delx, shiftyc = yc + dely;
    float cang = cosf(ang), sang = sinf(ang);
    for (int iy = 0; iy < ny; ++iy) {
        float y = (float)iy - shiftyc;
        float ycang = y * cang / 1.0f + yc;
        float ysang = -y * sang / 1.0f + xc;
        for (int ix = 0; ix < nx; ++ix) {
            float x = (float)ix - shiftxc;
            float xold = x * cang / 1.0f + ysang;
            float yold = x * sang / 1.0f + ycang;
            dst[ix + (size_t)iy * nx] =
                quadri_background(xold + 1.0f, yold + 1.0f, nx, ny, src, ix + 1, iy + 1);
        }
    }
    if (mirror) {
        int x_start = 1 - nx % 2;
        for (int iy = 0; iy < ny; ++iy) {
            float *row = dst + (size_t)iy * nx;
            for (int a = x_start, b = nx - 1; a < b; ++a, --b) { float t = row[a]; row[a] = row[b]; row[b] = t; }
        }
    }
}

/* ------------------------------------------------------------------ */
/* model_circle / normalize.mask (processor.cpp; test_mref.py:134,171,188) */

void cra_o_model_circle(float radius, int nx, int ny, float *mask)
{
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            float x2 = fabsf((float)i - nx / 2), y2 = fabsf((float)j - ny / 2);
            float r = (x2 * x2) / (radius * radius) + (y2 * y2) / (radius * radius);
            mask[i + (size_t)j * nx] = (r <= 1) ? 1.0f : 0.0f;
        }
}

/* no_sigma==0 -> subtract masked mean only; no_sigma==1 -> also divide by
 * the masked (n-1) standard deviation (yes, the flag reads backwards). */
void cra_o_normalize_mask(float *img, const float *mask, int n, int no_sigma)
{
    double sum = 0, sq2 = 0; size_t nn = 0;
    for (int i = 0; i < n; ++i)
        if (mask[i] > 0.5f) { sum += img[i]; sq2 += img[i] * (double)img[i]; ++nn; }
    float mean = (nn == 0) ? 0.0f : (float)sum / nn;
    float sigma = 1.0f;
    if (no_sigma != 0) sigma = sqrtf((float)((sq2 - sum * sum / nn) / (nn - 1)));
    for (int i = 0; i < n; ++i) img[i] = (img[i] - mean) / sigma;
}

/* search_range (sp_alignment.py) after the driver's swap (test_mref.py:195-198):
 * returns [left, right] */
void cra_o_search_range(int n, int radius, double shift, double range, double *lr)
{
    int cn = n / 2 + 1;
    double ql = cn + shift - radius - 2;
    double qe = n - cn - shift - radius;
    if (ql < 0) ql = 0;
    if (qe < 0) qe = 0;
    lr[0] = ql < range ? ql : range;
    lr[1] = qe < range ? qe : range;
}

/* ------------------------------------------------------------------ */
/* Reference preparation (test_mref.py:170-175): refs are normalised in
 * place (no_sigma=1), resampled at (cnx,cny), FFT'd and weighted.       */
void cra_o_prepare_refs(float *refs, int R, int nx, const float *mask,
                        const int *numr, int nring, float *crefim)
{
    int lcirc = numr[3 * nring - 2] + numr[3 * nring - 1] - 1;
    float wr[4096];
    cra_o_ringwe(numr, nring, wr);
    float cnx = (float)(nx / 2 + 1);
    for (int j = 0; j < R; ++j) {
        float *im = refs + (size_t)j * nx * nx;
        cra_o_normalize_mask(im, mask, nx * nx, 1);
        float *c = crefim + (size_t)j * lcirc;
        cra_o_polar2dm(im, nx, nx, cnx, cnx, numr, nring, c);
        cra_o_frngs(c, numr, nring);
        cra_o_applyws(c, numr, nring, wr);
    }
}

/* ------------------------------------------------------------------ */
/* One iteration of the per-particle section of mref_ali2d_MPI
 * (test_mref.py:183-215) for particles [0,P) whose global indices are
 * gofs..gofs+P-1.  params: [P][4] = alpha,sx,sy,mirror (double), updated in
 * place.  assign[P], peak[P].  sums: [R][2][nx][nx] (accumulated, caller
 * zeroes), counts[R] (accumulated).  normalize: 1 = multiref semantics.
 * Threads split particles; class sums are reduced in thread order so the
 * result does not depend on scheduling.                                 */
void cra_o_mref_iteration(float *images, int P, int nx, const float *mask,
                          const float *crefim, int R, const int *numr, int nring,
                          double xrng, double yrng, double step, int ou,
                          double *params, int *assign, float *peak,
                          float *sums, double *counts, long gofs, int normalize,
                          int nthreads)
{
    int cnx = nx / 2 + 1, cny = cnx;
    int mashi = cnx - ou - 2;
    size_t npix = (size_t)nx * nx;
    if (nthreads < 1) nthreads = 1;
    float *tsums = (float *)calloc((size_t)nthreads * R * 2 * npix, sizeof(float));
    double *tcnt = (double *)calloc((size_t)nthreads * R, sizeof(double));
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num(), nth = omp_get_num_threads();
#else
        int tid = 0, nth = 1;
#endif
        long lo = (long)P * tid / nth, hi = (long)P * (tid + 1) / nth;
        float *temp = (float *)malloc(sizeof(float) * npix);
        float *mys = tsums + (size_t)tid * R * 2 * npix;
        for (long im = lo; im < hi; ++im) {
            float *img = images + (size_t)im * npix;
            double *pp = params + 4 * im;
            double inv[4];
            cra_o_inverse_transform2(pp[0], pp[1], pp[2], 0, inv);
            double sxi = inv[1], syi = inv[2];
            cra_o_normalize_mask(img, mask, (int)npix, 0);
            if (fabs(sxi) > mashi || fabs(syi) > mashi) {
                sxi = 0.0; syi = 0.0;
                pp[0] = pp[1] = pp[2] = pp[3] = 0.0;
            }
            double tx[2], ty[2];
            cra_o_search_range(nx, ou, sxi, xrng, tx);
            cra_o_search_range(nx, ou, syi, yrng, ty);
            float res[8];
            cra_o_multiref_polar_ali_2d(img, nx, nx, crefim, R,
                                        (float)tx[0], (float)tx[1], (float)ty[0], (float)ty[1],
                                        (float)step, numr, nring,
                                        (float)(cnx + sxi), (float)(cny + syi), normalize, res);
            int iref = (int)res[4];
            double comb[4];
            cra_o_combine_params2(0.0, -sxi, -syi, 0, res[0], res[1], res[2], (int)res[3], comb);
            pp[0] = comb[0]; pp[1] = comb[1]; pp[2] = comb[2]; pp[3] = comb[3];
            assign[im] = iref; peak[im] = res[5];
            cra_o_rot_shift2d(img, nx, nx, (float)comb[0], (float)comb[1], (float)comb[2],
                              (int)comb[3], temp);
            int it = (int)((gofs + im) % 2);
            float *dst = mys + ((size_t)iref * 2 + it) * npix;
            for (size_t i = 0; i < npix; ++i) dst[i] += temp[i];
            tcnt[(size_t)tid * R + iref] += 1.0;
        }
        free(temp);
    }
    for (int t = 0; t < nthreads; ++t) {
        const float *s = tsums + (size_t)t * R * 2 * npix;
        for (size_t i = 0; i < (size_t)R * 2 * npix; ++i) sums[i] += s[i];
        for (int r = 0; r < R; ++r) counts[r] += tcnt[(size_t)t * R + r];
    }
    free(tsums); free(tcnt);
}

/* Alignment only (no class sums): used for timing the a7 hot spot and for
 * per-particle parity checks against explicit windows/centres.
 * centres: [P][2] float (cnx+sxi, cny+syi); win: [P][4] float xl,xr,yl,yr.
 * out: [P][8] as cra_o_multiref_polar_ali_2d.                           */
void cra_o_align_batch(const float *images, int P, int nx,
                       const float *crefim, int R, const int *numr, int nring,
                       const float *centres, const float *win, float step,
                       int normalize, float *out, int nthreads)
{
    size_t npix = (size_t)nx * nx;
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
    for (long im = 0; im < P; ++im)
        cra_o_multiref_polar_ali_2d(images + (size_t)im * npix, nx, nx, crefim, R,
                                    win[4 * im], win[4 * im + 1], win[4 * im + 2], win[4 * im + 3],
                                    step, numr, nring, centres[2 * im], centres[2 * im + 1],
                                    normalize, out + 8 * im);
}

int cra_o_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#ifdef __cplusplus
}
#endif
