/*
 * oracle/astra_linear.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic the reference delegates to the ASTRA toolbox
 * (un-vendored, un-pinned third-party dependency; era => astra-toolbox 1.8.x):
 *
 *   reference call site                      ASTRA algorithm / class restated here
 *   src/astra_proxy.py:35-49  gpu_fp    ->   'FP'   + 'linear' projector (CParallelBeamLinearKernelProjector2D)
 *   src/astra_proxy.py:52-66  gpu_bp    ->   'BP'   (exact transpose, scatter)
 *   src/astra_proxy.py:69-84  cpu_sirt  ->   'SIRT' (CSirtAlgorithm)
 *   src/astra_proxy.py:87-102 cpu_fbp   ->   'FBP'  (CFilteredBackProjectionAlgorithm, Fourier ramp)
 *
 * ASTRA is absent from /root/reference and cannot be installed here, so the published
 * algorithm is restated and pinned on the reference's own ASTRA-made
 * artifacts: teeth_recon/phantom.txt + teeth_recon/projjections.txt (G1) and src/fbp.nptxt (G2);
 * see tests/test_oracle_golden.py.  SIRT has no artifact in the tree: "parity unpinned" for SIRT.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (whitomo_b200/) never does.
 *
 * Geometry conventions: volume R x C, pixel size 1, centred; row 0 is the top
 * (+y up); angle theta stored as float32; ray direction (-sin, cos) (only ratios matter);
 * detector axis (cos, sin) * width; sinogram (A, D) C-order.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef struct {
    int rows, cols, ndet, nangles;
    float det_width;
    const float *angles;
} orc_geom;

/* per-angle constants exactly as the vector geometry would hold them (float32 fields) */
typedef struct {
    float ray_x, ray_y, det_sx, det_sy, det_ux, det_uy;
} par_proj;

static par_proj make_proj(const orc_geom *g, int a)
{
    par_proj p;
    double th = (double)g->angles[a];
    double c = cos(th), s = sin(th);
    p.ray_x = (float)(-s);
    p.ray_y = (float)(c);
    p.det_ux = (float)((double)g->det_width * c);
    p.det_uy = (float)((double)g->det_width * s);
    p.det_sx = (float)((double)g->det_width * -0.5 * (double)g->ndet * c);
    p.det_sy = (float)((double)g->det_width * -0.5 * (double)g->ndet * s);
    return p;
}

/* modes of the single ray walker (ASTRA's "policies") */
enum { MODE_FP = 0, MODE_BP = 1, MODE_RAYSUM = 2, MODE_PIXSUM = 3 };

/*
 * Walk every ray of angles [a0,a1) in ASTRA's order (angle, detector, row|col), float32
 * arithmetic with the running sum c += deltac, no FMA contraction
 * (compile with -ffp-contract=off).
 *   MODE_FP     : sino[ray] += vol[pix] * w
 *   MODE_BP     : vol[pix]  += sino[ray] * w
 *   MODE_RAYSUM : sino[ray] += w          (W 1, SIRT row sums)
 *   MODE_PIXSUM : vol[pix]  += w          (W^T 1, SIRT column sums)
 */
static void walk(const orc_geom *g, int a0, int a1, int mode, float *vol, float *sino)
{
    const int R = g->rows, C = g->cols, D = g->ndet;
    const float Ex = -0.5f * (float)C + 0.5f;
    const float Ey = 0.5f * (float)R - 0.5f;
    for (int a = a0; a < a1; ++a) {
        par_proj p = make_proj(g, a);
        const int vertical = fabsf(p.ray_x) < fabsf(p.ray_y);
        float ratio, len, delta;
        if (vertical) {
            ratio = p.ray_x / p.ray_y;
            len = sqrtf(p.ray_y * p.ray_y + p.ray_x * p.ray_x) / fabsf(p.ray_y);
            delta = -ratio;
        } else {
            ratio = p.ray_y / p.ray_x;
            len = sqrtf(p.ray_y * p.ray_y + p.ray_x * p.ray_x) / fabsf(p.ray_x);
            delta = -ratio;
        }
        for (int d = 0; d < D; ++d) {
            const int ray = a * D + d;
            const float Dx = p.det_sx + ((float)d + 0.5f) * p.det_ux;
            const float Dy = p.det_sy + ((float)d + 0.5f) * p.det_uy;
            int isin = 0;
            if (vertical) {
                float c = (Dx + (Ey - Dy) * ratio - Ex);
                for (int row = 0; row < R; ++row, c += delta) {
                    int col = (int)floorf(c);
                    if (col < -1 || col >= C) { if (!isin) continue; else break; }
                    float off = c - (float)col;
                    int pix = row * C + col;
                    if (col >= 0) {
                        float w = (1.0f - off) * len;
                        switch (mode) {
                        case MODE_FP: sino[ray] += vol[pix] * w; break;
                        case MODE_BP: vol[pix] += sino[ray] * w; break;
                        case MODE_RAYSUM: sino[ray] += w; break;
                        default: vol[pix] += w; break;
                        }
                    }
                    if (col + 1 < C) {
                        float w = off * len;
                        switch (mode) {
                        case MODE_FP: sino[ray] += vol[pix + 1] * w; break;
                        case MODE_BP: vol[pix + 1] += sino[ray] * w; break;
                        case MODE_RAYSUM: sino[ray] += w; break;
                        default: vol[pix + 1] += w; break;
                        }
                    }
                    isin = 1;
                }
            } else {
                float r = -(Dy + (Ex - Dx) * ratio - Ey);
                for (int col = 0; col < C; ++col, r += delta) {
                    int row = (int)floorf(r);
                    if (row < -1 || row >= R) { if (!isin) continue; else break; }
                    float off = r - (float)row;
                    int pix = row * C + col;
                    if (row >= 0) {
                        float w = (1.0f - off) * len;
                        switch (mode) {
                        case MODE_FP: sino[ray] += vol[pix] * w; break;
                        case MODE_BP: vol[pix] += sino[ray] * w; break;
                        case MODE_RAYSUM: sino[ray] += w; break;
                        default: vol[pix] += w; break;
                        }
                    }
                    if (row + 1 < R) {
                        float w = off * len;
                        switch (mode) {
                        case MODE_FP: sino[ray] += vol[pix + C] * w; break;
                        case MODE_BP: vol[pix + C] += sino[ray] * w; break;
                        case MODE_RAYSUM: sino[ray] += w; break;
                        default: vol[pix + C] += w; break;
                        }
                    }
                    isin = 1;
                }
            }
        }
    }
}

static orc_geom mk(int rows, int cols, int ndet, float detw, const float *angles, int nang)
{
    orc_geom g;
    g.rows = rows; g.cols = cols; g.ndet = ndet; g.nangles = nang;
    g.det_width = detw; g.angles = angles;
    return g;
}

/* ---- faithful single-thread float32 operators (what ASTRA's CPU algorithms run) ---- */

int orc_fp(int rows, int cols, int ndet, float detw, const float *angles, int nang,
           const float *vol, float *sino)
{
    orc_geom g = mk(rows, cols, ndet, detw, angles, nang);
    memset(sino, 0, sizeof(float) * (size_t)nang * ndet);
    walk(&g, 0, nang, MODE_FP, (float *)vol, sino);
    return 0;
}

int orc_bp(int rows, int cols, int ndet, float detw, const float *angles, int nang,
           const float *sino, float *vol)
{
    orc_geom g = mk(rows, cols, ndet, detw, angles, nang);
    memset(vol, 0, sizeof(float) * (size_t)rows * cols);
    walk(&g, 0, nang, MODE_BP, vol, (float *)sino);
    return 0;
}

/* ---- all-host-cores variants (pthreads over angle blocks; same arithmetic per ray;
 *      BP reduces per-thread partial volumes, so its summation order differs from orc_bp) ---- */

typedef struct {
    const orc_geom *g;
    int mode, tid, nth;
    float *vol, *sino;
} mt_job;

static void *mt_run(void *arg)
{
    mt_job *j = (mt_job *)arg;
    /* interleaved blocks of <= 4 angles balance the vertical/horizontal halves across threads */
    int blk = j->g->nangles / (j->nth * 4);
    if (blk < 1) blk = 1;
    if (blk > 4) blk = 4;
    for (int a = j->tid * blk; a < j->g->nangles; a += j->nth * blk) {
        int a1 = a + blk < j->g->nangles ? a + blk : j->g->nangles;
        walk(j->g, a, a1, j->mode, j->vol, j->sino);
    }
    return NULL;
}

static int mt_threads(int req)
{
    long n = req > 0 ? req : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}

int orc_num_threads(void) { return mt_threads(0); }

int orc_fp_mt(int rows, int cols, int ndet, float detw, const float *angles, int nang,
              const float *vol, float *sino, int nthreads)
{
    orc_geom g = mk(rows, cols, ndet, detw, angles, nang);
    int nth = mt_threads(nthreads);
    memset(sino, 0, sizeof(float) * (size_t)nang * ndet);
    pthread_t th[256];
    mt_job jobs[256];
    for (int t = 0; t < nth; ++t) {
        jobs[t].g = &g; jobs[t].mode = MODE_FP; jobs[t].tid = t; jobs[t].nth = nth;
        jobs[t].vol = (float *)vol; jobs[t].sino = sino;
        if (pthread_create(&th[t], NULL, mt_run, &jobs[t])) return -2;
    }
    for (int t = 0; t < nth; ++t) pthread_join(th[t], NULL);
    return 0;
}

int orc_bp_mt(int rows, int cols, int ndet, float detw, const float *angles, int nang,
              const float *sino, float *vol, int nthreads)
{
    orc_geom g = mk(rows, cols, ndet, detw, angles, nang);
    int nth = mt_threads(nthreads);
    const size_t npix = (size_t)rows * cols;
    memset(vol, 0, sizeof(float) * npix);
    float *part = (float *)calloc(npix * (size_t)nth, sizeof(float));
    if (!part) return -1;
    pthread_t th[256];
    mt_job jobs[256];
    for (int t = 0; t < nth; ++t) {
        jobs[t].g = &g; jobs[t].mode = MODE_BP; jobs[t].tid = t; jobs[t].nth = nth;
        jobs[t].vol = part + npix * (size_t)t; jobs[t].sino = (float *)sino;
        if (pthread_create(&th[t], NULL, mt_run, &jobs[t])) return -2;
    }
    for (int t = 0; t < nth; ++t) pthread_join(th[t], NULL);
    for (int t = 0; t < nth; ++t) {
        const float *mine = part + npix * (size_t)t;
        for (size_t i = 0; i < npix; ++i) vol[i] += mine[i];
    }
    free(part);
    return 0;
}

/* ---- float64 closed-form arbiter (SURVEY A.2 / A.3 formulas, no running sums) ---- */

static void fp_f64_range(int rows, int cols, int ndet, float detw, const float *angles, int a0, int a1,
                         const double *vol, double *sino)
{
    const int R = rows, C = cols, D = ndet;
    for (int a = a0; a < a1; ++a) {
        float thf = angles[a];
        double th = (double)thf;
        double cs = cos(th), sn = sin(th);
        /* the branch predicate is evaluated on the float32 ray vector, like ASTRA */
        int vertical = fabsf((float)(-sn)) < fabsf((float)cs);
        for (int d = 0; d < D; ++d) {
            double s = ((double)d + 0.5 - 0.5 * D) * (double)detw;
            double acc = 0.0;
            if (vertical) {
                double L = 1.0 / fabs(cs), tn = sn / cs;
                for (int row = 0; row < R; ++row) {
                    double y = 0.5 * R - 0.5 - row;
                    double c = s / cs - y * tn + 0.5 * C - 0.5;
                    double fl = floor(c);
                    int col = (int)fl;
                    if (col < -1 || col >= C) continue;
                    double off = c - fl;
                    if (col >= 0) acc += vol[(size_t)row * C + col] * (1.0 - off) * L;
                    if (col + 1 < C) acc += vol[(size_t)row * C + col + 1] * off * L;
                }
            } else {
                double L = 1.0 / fabs(sn), ct = cs / sn;
                for (int col = 0; col < C; ++col) {
                    double x = col - 0.5 * C + 0.5;
                    double r = (0.5 * R - 0.5) - s / sn + x * ct;
                    double fl = floor(r);
                    int row = (int)fl;
                    if (row < -1 || row >= R) continue;
                    double off = r - fl;
                    if (row >= 0) acc += vol[(size_t)row * C + col] * (1.0 - off) * L;
                    if (row + 1 < R) acc += vol[(size_t)(row + 1) * C + col] * off * L;
                }
            }
            sino[(size_t)a * D + d] = acc;
        }
    }
}

int orc_fp_f64(int rows, int cols, int ndet, float detw, const float *angles, int nang,
               const double *vol, double *sino)
{
    fp_f64_range(rows, cols, ndet, detw, angles, 0, nang, vol, sino);
    return 0;
}

/* pixel-driven gather form of the exact transpose, float64 */
static void bp_f64_rows(int rows, int cols, int ndet, float detw, const float *angles, int nang,
                        const double *sino, double *vol, int r0, int r1)
{
    const int R = rows, C = cols, D = ndet;
    memset(vol + (size_t)r0 * C, 0, sizeof(double) * (size_t)(r1 - r0) * C);
    for (int a = 0; a < nang; ++a) {
        double th = (double)angles[a];
        double cs = cos(th), sn = sin(th);
        int vertical = fabsf((float)(-sn)) < fabsf((float)cs);
        double m = vertical ? fabs(cs) : fabs(sn);
        const double *q = sino + (size_t)a * D;
        for (int row = r0; row < r1; ++row) {
            double y = 0.5 * R - 0.5 - row;
            for (int col = 0; col < C; ++col) {
                double x = col - 0.5 * C + 0.5;
                double ds = (x * cs + y * sn) / (double)detw + 0.5 * D - 0.5;
                double fl = floor(ds);
                int d0 = (int)fl;
                double f = ds - fl;
                double acc = 0.0;
                /* detector spacing detw: the footprint half-width in detector units is m/detw */
                double hw = m / (double)detw;
                if (d0 >= 0 && d0 < D) {
                    double w = 1.0 - f / hw;
                    if (w > 0) acc += w / m * q[d0];
                }
                if (d0 + 1 >= 0 && d0 + 1 < D) {
                    double w = 1.0 - (1.0 - f) / hw;
                    if (w > 0) acc += w / m * q[d0 + 1];
                }
                vol[(size_t)row * C + col] += acc;
            }
        }
    }
}

int orc_bp_f64(int rows, int cols, int ndet, float detw, const float *angles, int nang,
               const double *sino, double *vol)
{
    bp_f64_rows(rows, cols, ndet, detw, angles, nang, sino, vol, 0, rows);
    return 0;
}

/* all-host-cores variants of the float64 arbiter: FP over angle ranges, BP over row ranges (same arithmetic and
 * summation order per output element as the single-thread functions) */
typedef struct {
    int rows, cols, ndet, nang, lo, hi, is_bp;
    float detw;
    const float *angles;
    const double *in;
    double *out;
} f64_job;

static void *f64_run(void *arg)
{
    f64_job *j = (f64_job *)arg;
    if (j->is_bp) bp_f64_rows(j->rows, j->cols, j->ndet, j->detw, j->angles, j->nang, j->in, j->out, j->lo, j->hi);
    else fp_f64_range(j->rows, j->cols, j->ndet, j->detw, j->angles, j->lo, j->hi, j->in, j->out);
    return NULL;
}

static int f64_mt(int rows, int cols, int ndet, float detw, const float *angles, int nang,
                  const double *in, double *out, int nthreads, int is_bp)
{
    int nth = mt_threads(nthreads);
    const int total = is_bp ? rows : nang;
    if (nth > total) nth = total;
    pthread_t th[256];
    f64_job jobs[256];
    for (int t = 0; t < nth; ++t) {
        f64_job *j = &jobs[t];
        j->rows = rows; j->cols = cols; j->ndet = ndet; j->nang = nang; j->detw = detw; j->angles = angles;
        j->in = in; j->out = out; j->is_bp = is_bp;
        j->lo = (int)((long)total * t / nth); j->hi = (int)((long)total * (t + 1) / nth);
        if (pthread_create(&th[t], NULL, f64_run, j)) return -2;
    }
    for (int t = 0; t < nth; ++t) pthread_join(th[t], NULL);
    return 0;
}

int orc_fp_f64_mt(int rows, int cols, int ndet, float detw, const float *angles, int nang,
                  const double *vol, double *sino, int nthreads)
{
    return f64_mt(rows, cols, ndet, detw, angles, nang, vol, sino, nthreads, 0);
}

int orc_bp_f64_mt(int rows, int cols, int ndet, float detw, const float *angles, int nang,
                  const double *sino, double *vol, int nthreads)
{
    return f64_mt(rows, cols, ndet, detw, angles, nang, sino, vol, nthreads, 1);
}

/* ---- SIRT (ASTRA CSirtAlgorithm, lambda = 1, no constraints). UNPINNED. ---- */

static void guarded_div(float *num, const float *den, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        if (den[i] > 0.000001f) num[i] /= den[i];
        else num[i] = 0.0f;
    }
}

int orc_sirt(int rows, int cols, int ndet, float detw, const float *angles, int nang,
             const float *sino, float *vol, int iters)
{
    orc_geom g = mk(rows, cols, ndet, detw, angles, nang);
    const size_t npix = (size_t)rows * cols, nray = (size_t)nang * ndet;
    float *raysum = (float *)calloc(nray, sizeof(float));
    float *pixsum = (float *)calloc(npix, sizeof(float));
    float *diff = (float *)malloc(nray * sizeof(float));
    float *tmp = (float *)malloc(npix * sizeof(float));
    if (!raysum || !pixsum || !diff || !tmp) return -1;
    walk(&g, 0, nang, MODE_RAYSUM, NULL, raysum);
    walk(&g, 0, nang, MODE_PIXSUM, pixsum, NULL);
    memset(vol, 0, npix * sizeof(float));
    for (int it = 0; it < iters; ++it) {
        /* diff = sino - W x   (ASTRA DiffFPPolicy: diff starts as sino, subtracts per tap) */
        memset(diff, 0, nray * sizeof(float));
        walk(&g, 0, nang, MODE_FP, vol, diff);
        for (size_t i = 0; i < nray; ++i) diff[i] = sino[i] - diff[i];
        guarded_div(diff, raysum, nray);
        memset(tmp, 0, npix * sizeof(float));
        walk(&g, 0, nang, MODE_BP, tmp, diff);
        guarded_div(tmp, pixsum, npix);
        for (size_t i = 0; i < npix; ++i) vol[i] += tmp[i];
    }
    free(raysum); free(pixsum); free(diff); free(tmp);
    return 0;
}

/* ---- FBP (verified against src/fbp.nptxt) ---- */

static void fft_inplace(double *re, double *im, int n, int inverse)
{
    /* iterative radix-2 */
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        double ang = 2.0 * M_PI / len * (inverse ? 1.0 : -1.0);
        double wr = cos(ang), wi = sin(ang);
        for (int i = 0; i < n; i += len) {
            double cr = 1.0, ci = 0.0;
            for (int k = 0; k < len / 2; ++k) {
                double ur = re[i + k], ui = im[i + k];
                double vr = re[i + k + len / 2] * cr - im[i + k + len / 2] * ci;
                double vi = re[i + k + len / 2] * ci + im[i + k + len / 2] * cr;
                re[i + k] = ur + vr; im[i + k] = ui + vi;
                re[i + k + len / 2] = ur - vr; im[i + k + len / 2] = ui - vi;
                double ncr = cr * wr - ci * wi;
                ci = cr * wi + ci * wr; cr = ncr;
            }
        }
    }
    if (inverse)
        for (int i = 0; i < n; ++i) { re[i] /= n; im[i] /= n; }
}

/* ramp-filter every sinogram row: zero-pad to zp = pow2 >= 2D, multiply bin i by 2*min(i, zp-i)/zp */
int orc_ramp_filter(int ndet, int nang, const float *sino, float *out)
{
    int zp = 1;
    while (zp < 2 * ndet) zp <<= 1;
    double *re = (double *)malloc(sizeof(double) * zp), *im = (double *)malloc(sizeof(double) * zp);
    if (!re || !im) return -1;
    for (int a = 0; a < nang; ++a) {
        for (int i = 0; i < zp; ++i) { re[i] = i < ndet ? (double)sino[(size_t)a * ndet + i] : 0.0; im[i] = 0.0; }
        fft_inplace(re, im, zp, 0);
        for (int i = 0; i < zp; ++i) {
            int k = i <= zp / 2 ? i : zp - i;
            double f = (double)((2.0f * (float)k) / (float)zp);
            re[i] *= f; im[i] *= f;
        }
        fft_inplace(re, im, zp, 1);
        for (int i = 0; i < ndet; ++i) out[(size_t)a * ndet + i] = (float)re[i];
    }
    free(re); free(im);
    return 0;
}

int orc_fbp(int rows, int cols, int ndet, float detw, const float *angles, int nang,
            const float *sino, float *vol)
{
    float *filt = (float *)malloc(sizeof(float) * (size_t)nang * ndet);
    if (!filt) return -1;
    orc_ramp_filter(ndet, nang, sino, filt);
    orc_bp(rows, cols, ndet, detw, angles, nang, filt, vol);
    const float scale = (float)(M_PI / 2.0) / (float)nang;
    for (size_t i = 0; i < (size_t)rows * cols; ++i) vol[i] *= scale;
    free(filt);
    return 0;
}
