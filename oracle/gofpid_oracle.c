/*
 * gofpid_oracle.c -- CPU restatement of PyGOFPID's per-frame foreground-extraction
 * hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  It is the checker for the CUDA path in
 * pygofpid_b200/csrc, never the thing that is shipped or measured as the product.
 *
 * The reference (qbarthelemy/PyGOFPID v1.1.0) is pure Python and delegates every
 * arithmetic step to OpenCV (third-party, un-vendored; reference pins
 * opencv-python>=3.4,<4 in setup.py:14, the image carries opencv-python-headless
 * 4.13.0).  Each function below restates the published OpenCV algorithm behind one
 * reference call site and cites that call site.  Parity pin: tests/test_oracle_vs_cv2.py
 * checks every function here bit-for-bit against the installed cv2 (live), and
 * tests/golden/ holds vectors produced by running the real reference package
 * (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).  -ffp-contract=off
 * matters: the MOG2 arithmetic must not be fused into FMAs (cv2's video module is
 * built for baseline SSE3).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* border handling: BORDER_REFLECT_101 (cv2 BORDER_DEFAULT), iterated for tiny images */
static int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

/* ------------------------------------------------------------------------- */
/* Gaussian kernel in 8.8 fixed point.
 * Reference call site: pipeline.py:463-467 -> cv.GaussianBlur(X, ksize, sigmaX=borderType).
 * cv2 builds a float64 kernel exp(-(i-c)^2/(2 s^2)) (normalised), then quantises the first
 * half with error diffusion and sets the centre so that the taps sum to exactly 256. */
ORC_API int orc_gauss_kernel_q8(int n, double sigma, uint16_t *q)
{
    if (n <= 0 || (n & 1) == 0 || n > 255) return -1;
    double k[256];
    if (n == 1) { q[0] = 256; return 0; }
    if (sigma <= 0 && n <= 9) {
        static const double t3[] = {0.25, 0.5, 0.25};
        static const double t5[] = {0.0625, 0.25, 0.375, 0.25, 0.0625};
        static const double t7[] = {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125};
        static const double t9[] = {4 / 256., 13 / 256., 30 / 256., 51 / 256., 60 / 256.,
                                    51 / 256., 30 / 256., 13 / 256., 4 / 256.};
        const double *t = n == 3 ? t3 : n == 5 ? t5 : n == 7 ? t7 : t9;
        for (int i = 0; i < n; i++) k[i] = t[i];
    } else {
        double s = sigma > 0 ? sigma : 0.15 * n + 0.35; /* = 0.3*((n-1)*0.5-1)+0.8 */
        double scale2x = -0.125 / (s * s);               /* x below is doubled */
        int h = (n - 1) / 2;
        double sum = 0;
        for (int i = 0, x = 1 - n; i < h; i++, x += 2) {
            k[i] = exp((double)(x * x) * scale2x);
            sum += k[i];
        }
        sum = sum * 2 + 1;
        double mul = 1.0 / sum;
        for (int i = 0; i < h; i++) { k[i] *= mul; k[n - 1 - i] = k[i]; }
        k[h] = mul;
    }
    int h = n / 2;
    double err = 0;
    long acc = 0;
    for (int i = 0; i < h; i++) {
        double adj = k[i] * 256.0 + err;
        long v = lrint(adj); /* round half to even, as cvRound */
        err = adj - (double)v;
        q[i] = q[n - 1 - i] = (uint16_t)v;
        acc += v;
    }
    q[h] = (uint16_t)(256 - 2 * acc);
    return 0;
}

/* Separable fixed-point blur, u8 -> u8, interleaved channels.
 * Row pass: u16 = sum q*src (fits: taps sum to 256).  Column pass: u32 = sum q*row.
 * Output = (v + 2^15) >> 16.  Border: REFLECT_101 on both axes. */
ORC_API void orc_blur_q8(const uint8_t *src, int H, int W, int C,
                         const uint16_t *qx, int nx, const uint16_t *qy, int ny, uint8_t *dst)
{
    int rx = nx / 2, ry = ny / 2;
    uint16_t *row = (uint16_t *)malloc((size_t)H * W * C * sizeof(uint16_t));
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            for (int c = 0; c < C; c++) {
                uint32_t s = 0;
                for (int i = 0; i < nx; i++)
                    s += (uint32_t)qx[i] * src[((size_t)y * W + reflect101(x + i - rx, W)) * C + c];
                row[((size_t)y * W + x) * C + c] = (uint16_t)s;
            }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            for (int c = 0; c < C; c++) {
                uint32_t s = 0;
                for (int i = 0; i < ny; i++)
                    s += (uint32_t)qy[i] * row[((size_t)reflect101(y + i - ry, H) * W + x) * C + c];
                uint32_t v = (s + 32768u) >> 16;
                dst[((size_t)y * W + x) * C + c] = (uint8_t)(v > 255 ? 255 : v);
            }
    free(row);
}

/* Normalised box filter.  Reference call site: pipeline.py:463 with fun=cv.blur
 * (tests/test_pipeline.py:68-75).  REFLECT_101, anchor k/2.  For k-areas up to 256 cv2 keeps
 * u16 column sums and divides by multiplication: out = ((S + delta) * scale) >> 23 with
 * scale = floor(2^23/d) (+1 when the fraction is >= .5) and delta = d/2 (+1 otherwise); this
 * is round-to-nearest for odd d and has its own tie behaviour for even d.  Larger areas use
 * a float scale: out = rint(S * (1/d)). */
ORC_API void orc_box_divisor(int d, int *delta, int *scale)
{
    double sf = (double)(1 << 23) / d;
    int sc = (int)floor(sf);
    int dl = d / 2;
    if (sf - sc < 0.5) dl++; else sc++;
    *delta = dl; *scale = sc;
}

ORC_API void orc_box_u8(const uint8_t *src, int H, int W, int C, int kx, int ky, uint8_t *dst)
{
    int ax = kx / 2, ay = ky / 2;
    int d = kx * ky, delta, scale;
    orc_box_divisor(d, &delta, &scale);
    double fscale = 1.0 / (double)d;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            for (int c = 0; c < C; c++) {
                int s = 0;
                for (int j = 0; j < ky; j++) {
                    int yy = reflect101(y + j - ay, H);
                    for (int i = 0; i < kx; i++)
                        s += src[((size_t)yy * W + reflect101(x + i - ax, W)) * C + c];
                }
                long v;
                if (d == 1) v = s;
                else if (d <= 256) v = (long)(((int64_t)(s + delta) * scale) >> 23);
                else v = lrint(s * fscale);
                dst[((size_t)y * W + x) * C + c] = (uint8_t)(v > 255 ? 255 : v);
            }
}

/* ------------------------------------------------------------------------- */
/* Frame differencing.  Reference: segmentation.py:40-70 (absdiff :58, threshold :59-64,
 * uint8-accumulating mean over channels :67-68).  thr is already floor(threshold). */
ORC_API void orc_fd(const uint8_t *cur, const uint8_t *prev, int H, int W, int C, int thr,
                    uint8_t *mask)
{
    for (size_t p = 0; p < (size_t)H * W; p++) {
        if (C == 1) {
            int d = abs((int)cur[p] - (int)prev[p]);
            mask[p] = d > thr ? 255 : 0;
        } else {
            uint8_t acc = 0; /* np.mean(dtype=uint8): the sum wraps modulo 256 */
            for (int c = 0; c < C; c++) {
                int d = abs((int)cur[p * C + c] - (int)prev[p * C + c]);
                acc = (uint8_t)(acc + (d > thr ? 255 : 0));
            }
            mask[p] = (uint8_t)(acc / C);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* MOG2.  Reference call sites: pipeline.py:232 (cv.createBackgroundSubtractorMOG2(),
 * all defaults) and pipeline.py:470 (.apply(X)).  Zivkovic's adaptive GMM as published
 * in OpenCV's BackgroundSubtractorMOG2 (video module). */
#define MOG2_NMIX 5
typedef struct {
    int H, W, C;
    long nframes;
    float *weight;   /* [H*W][5] */
    float *var;      /* [H*W][5] */
    float *mean;     /* [H*W][5][C] */
    uint8_t *nmodes; /* [H*W] */
    /* parameters (OpenCV defaults) */
    int history;
    float Tb, TB, Tg, varInit, varMin, varMax, CT, tau;
    uint8_t shadowVal;
    int detectShadows;
} orc_mog2;

ORC_API orc_mog2 *orc_mog2_create(int H, int W, int C)
{
    orc_mog2 *m = (orc_mog2 *)calloc(1, sizeof(orc_mog2));
    size_t n = (size_t)H * W;
    m->H = H; m->W = W; m->C = C;
    m->weight = (float *)calloc(n * MOG2_NMIX, sizeof(float));
    m->var = (float *)calloc(n * MOG2_NMIX, sizeof(float));
    m->mean = (float *)calloc(n * MOG2_NMIX * C, sizeof(float));
    m->nmodes = (uint8_t *)calloc(n, 1);
    m->history = 500;
    m->Tb = 16.0f; m->TB = 0.9f; m->Tg = 9.0f;
    m->varInit = 15.0f; m->varMin = 4.0f; m->varMax = 75.0f;
    m->CT = 0.05f; m->tau = 0.5f;
    m->shadowVal = 127; m->detectShadows = 1;
    return m;
}

ORC_API void orc_mog2_destroy(orc_mog2 *m)
{
    if (!m) return;
    free(m->weight); free(m->var); free(m->mean); free(m->nmodes); free(m);
}

ORC_API float *orc_mog2_weight(orc_mog2 *m) { return m->weight; }
ORC_API float *orc_mog2_var(orc_mog2 *m) { return m->var; }
ORC_API float *orc_mog2_mean(orc_mog2 *m) { return m->mean; }
ORC_API uint8_t *orc_mog2_nmodes(orc_mog2 *m) { return m->nmodes; }
ORC_API long orc_mog2_nframes(orc_mog2 *m) { return m->nframes; }

static int mog2_shadow(const orc_mog2 *m, const float *x, int nmodes, const float *w,
                       const float *var, const float *mean)
{
    int C = m->C;
    float tW = 0.0f;
    for (int k = 0; k < nmodes; k++) {
        const float *mu = mean + k * C;
        float num = 0.0f, den = 0.0f;
        for (int c = 0; c < C; c++) {
            num += x[c] * mu[c];
            den += mu[c] * mu[c];
        }
        if (den == 0) return 0;
        if (num <= den && num >= m->tau * den) {
            float a = num / den;
            float d2 = 0.0f;
            for (int c = 0; c < C; c++) {
                float dd = a * mu[c] - x[c];
                d2 += dd * dd;
            }
            if (d2 < m->Tb * var[k] * a * a) return 1;
        }
        tW += w[k];
        if (tW > m->TB) return 0;
    }
    return 0;
}

ORC_API void orc_mog2_apply(orc_mog2 *m, const uint8_t *frame, uint8_t *mask)
{
    int C = m->C;
    m->nframes++;
    long n2 = 2 * m->nframes;
    double lr = 1.0 / (double)(n2 < m->history ? n2 : m->history);
    float alphaT = (float)lr;
    float alpha1 = 1.0f - alphaT;
    float prune = (float)(-lr * (double)m->CT);
    size_t npx = (size_t)m->H * m->W;

    for (size_t p = 0; p < npx; p++) {
        float x[4], d[4];
        for (int c = 0; c < C; c++) x[c] = (float)frame[p * C + c];
        float *w = m->weight + p * MOG2_NMIX;
        float *var = m->var + p * MOG2_NMIX;
        float *mean = m->mean + p * MOG2_NMIX * C;
        int nmodes = m->nmodes[p];
        int background = 0, fits = 0;
        float totalWeight = 0.0f;

        /* NB: the loop bound is the live mode count: a pruned mode shortens the walk. */
        for (int k = 0; k < nmodes; k++) {
            float weight = alpha1 * w[k] + prune;
            int swaps = 0;
            if (!fits) {
                float v = var[k];
                float *mu = mean + k * C;
                float dist2;
                if (C == 3) {
                    d[0] = mu[0] - x[0]; d[1] = mu[1] - x[1]; d[2] = mu[2] - x[2];
                    dist2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                } else {
                    dist2 = 0.0f;
                    for (int c = 0; c < C; c++) { d[c] = mu[c] - x[c]; dist2 += d[c] * d[c]; }
                }
                if (totalWeight < m->TB && dist2 < m->Tb * v) background = 1;
                if (dist2 < m->Tg * v) {
                    fits = 1;
                    weight += alphaT;
                    float kk = alphaT / weight;
                    for (int c = 0; c < C; c++) mu[c] -= kk * d[c];
                    float vn = v + kk * (dist2 - v);
                    vn = vn > m->varMin ? vn : m->varMin;
                    vn = vn < m->varMax ? vn : m->varMax;
                    var[k] = vn;
                    for (int i = k; i > 0; i--) {
                        if (weight < w[i - 1]) break;
                        swaps++;
                        float t;
                        t = w[i]; w[i] = w[i - 1]; w[i - 1] = t;
                        t = var[i]; var[i] = var[i - 1]; var[i - 1] = t;
                        for (int c = 0; c < C; c++) {
                            t = mean[i * C + c]; mean[i * C + c] = mean[(i - 1) * C + c];
                            mean[(i - 1) * C + c] = t;
                        }
                    }
                }
            }
            if (weight < -prune) { weight = 0.0f; nmodes--; }
            w[k - swaps] = weight;
            totalWeight += weight;
        }

        float inv = 0.0f;
        if (fabsf(totalWeight) > FLT_EPSILON) inv = 1.0f / totalWeight;
        for (int k = 0; k < nmodes; k++) w[k] *= inv;

        if (!fits && alphaT > 0.0f) {
            int k = (nmodes == MOG2_NMIX) ? MOG2_NMIX - 1 : nmodes++;
            if (nmodes == 1) w[k] = 1.0f;
            else {
                w[k] = alphaT;
                for (int i = 0; i < nmodes - 1; i++) w[i] *= alpha1;
            }
            for (int c = 0; c < C; c++) mean[k * C + c] = x[c];
            var[k] = m->varInit;
            for (int i = nmodes - 1; i > 0; i--) {
                if (alphaT < w[i - 1]) break;
                float t;
                t = w[i]; w[i] = w[i - 1]; w[i - 1] = t;
                t = var[i]; var[i] = var[i - 1]; var[i - 1] = t;
                for (int c = 0; c < C; c++) {
                    t = mean[i * C + c]; mean[i * C + c] = mean[(i - 1) * C + c];
                    mean[(i - 1) * C + c] = t;
                }
            }
        }
        m->nmodes[p] = (uint8_t)nmodes;
        mask[p] = background ? 0
                  : (m->detectShadows && mog2_shadow(m, x, nmodes, w, var, mean)) ? m->shadowVal
                                                                                   : 255;
    }
}

/* cv2's getBackgroundImage(): the only view of the MOG2 model cv2 exposes.  Weighted mean of
 * the leading modes until their cumulative weight exceeds backgroundRatio. */
ORC_API void orc_mog2_background(const orc_mog2 *m, uint8_t *out)
{
    int C = m->C;
    size_t npx = (size_t)m->H * m->W;
    for (size_t p = 0; p < npx; p++) {
        const float *w = m->weight + p * MOG2_NMIX;
        const float *mean = m->mean + p * MOG2_NMIX * C;
        float acc[4] = {0, 0, 0, 0};
        float tw = 0.0f;
        for (int k = 0; k < m->nmodes[p]; k++) {
            for (int c = 0; c < C; c++) acc[c] += w[k] * mean[k * C + c];
            tw += w[k];
            if (tw > m->TB) break;
        }
        float inv = tw != 0.0f ? 1.0f / tw : 0.0f;
        for (int c = 0; c < C; c++) {
            long v = lrintf(acc[c] * inv);
            out[p * C + c] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* cv::RNG replica (multiply-with-carry) and its uniform u8 fill.
 * KNN consumes the calling thread's cv::theRNG() through randu() on H*W u8 arrays. */
#define ORC_RNG_A 4164903690ULL
static inline uint32_t rng_next(uint64_t *s)
{
    *s = (uint64_t)(uint32_t)(*s) * ORC_RNG_A + (*s >> 32);
    return (uint32_t)(*s);
}

/* fill arr[0..total) with uniform integers in [0, range); advances *state exactly as
 * cv::RNG::fill(UNIFORM) does for a continuous CV_8UC1 matrix (blocks of 1024 elements). */
ORC_API void orc_rng_fill_u8(uint64_t *state, uint8_t *arr, long total, int range)
{
    uint32_t mask = (uint32_t)(range - 1);
    int pow2 = (range & (range - 1)) == 0;
    for (long j = 0; j < total; j += 1024) {
        long len = total - j < 1024 ? total - j : 1024;
        uint8_t *a = arr + j;
        long i = 0;
        if (pow2) {
            for (; i <= len - 4; i += 4) {
                uint32_t t = rng_next(state);
                a[i] = (uint8_t)(t & mask);
                a[i + 1] = (uint8_t)((t >> 8) & mask);
                a[i + 2] = (uint8_t)((t >> 16) & mask);
                a[i + 3] = (uint8_t)((t >> 24) & mask);
            }
            for (; i < len; i++) a[i] = (uint8_t)(rng_next(state) & mask);
        } else {
            for (; i < len; i++) a[i] = (uint8_t)(rng_next(state) % (uint32_t)range);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* KNN.  Reference call sites: pipeline.py:234 (cv.createBackgroundSubtractorKNN(),
 * defaults) and pipeline.py:470.  Zivkovic & van der Heijden's sample-based model. */
#define KNN_N 7
typedef struct {
    int H, W, C;
    long nframes;
    uint8_t *model;            /* [H*W][21][C+1] : tiers short | mid | long */
    uint8_t *idx[3];           /* ring positions S, M, L */
    uint8_t *nxt[3];           /* next-update tick S, M, L */
    int counter[3];            /* S, M, L */
    uint64_t rng;
    int history, knn;
    float Tb, tau;
    uint8_t shadowVal;
    uint8_t *fresh[3];         /* scratch for the per-tier refresh arrays */
} orc_knn;

ORC_API orc_knn *orc_knn_create(int H, int W, int C, uint64_t seed)
{
    orc_knn *k = (orc_knn *)calloc(1, sizeof(orc_knn));
    size_t n = (size_t)H * W;
    k->H = H; k->W = W; k->C = C;
    k->model = (uint8_t *)calloc(n * 3 * KNN_N * (C + 1), 1);
    for (int t = 0; t < 3; t++) {
        k->idx[t] = (uint8_t *)calloc(n, 1);
        k->nxt[t] = (uint8_t *)calloc(n, 1);
        k->fresh[t] = (uint8_t *)calloc(n, 1);
    }
    k->rng = seed ? seed : 0xffffffffULL;
    k->history = 500; k->knn = 2; k->Tb = 400.0f; k->tau = 0.5f; k->shadowVal = 127;
    return k;
}

ORC_API void orc_knn_destroy(orc_knn *k)
{
    if (!k) return;
    free(k->model);
    for (int t = 0; t < 3; t++) { free(k->idx[t]); free(k->nxt[t]); free(k->fresh[t]); }
    free(k);
}

ORC_API uint8_t *orc_knn_model(orc_knn *k) { return k->model; }
ORC_API uint8_t *orc_knn_idx(orc_knn *k, int t) { return k->idx[t]; }
ORC_API uint8_t *orc_knn_nxt(orc_knn *k, int t) { return k->nxt[t]; }
ORC_API uint64_t orc_knn_rng(orc_knn *k) { return k->rng; }
ORC_API int orc_knn_counter(orc_knn *k, int t) { return k->counter[t]; }

/* update periods (uS,uM,uL) for frame number n (1-based) */
ORC_API void orc_knn_periods(long n, int history, int *u)
{
    long n2 = 2 * n;
    double lr = 1.0 / (double)(n2 < history ? n2 : history);
    int Ks = (int)(log(0.7) / log(1 - lr)) + 1;
    int Km = (int)(log(0.4) / log(1 - lr)) - Ks + 1;
    int Kl = (int)(log(0.1) / log(1 - lr)) - Ks - Km + 1;
    u[0] = Ks / KNN_N + 1;
    u[1] = Km / KNN_N + 1;
    u[2] = Kl / KNN_N + 1;
}

static int knn_classify(const orc_knn *k, const uint8_t *x, const uint8_t *mdl, uint8_t *include)
{
    int C = k->C, nd = C + 1;
    int Pbf = 0, Pb = 0;
    *include = 0;
    for (int s = 0; s < 3 * KNN_N; s++) {
        const uint8_t *sm = mdl + s * nd;
        float dist2 = 0.0f;
        for (int c = 0; c < C; c++) {
            float d = (float)sm[c] - (float)x[c];
            dist2 += d * d;
        }
        if (dist2 < k->Tb) {
            Pbf++;
            if (sm[C]) {
                Pb++;
                if (Pb >= k->knn) { *include = 1; return 1; }
            }
        }
    }
    if (Pbf >= k->knn) *include = 1;
    int Ps = 0;
    for (int s = 0; s < 3 * KNN_N; s++) {
        const uint8_t *sm = mdl + s * nd;
        if (!sm[C]) continue;
        float num = 0.0f, den = 0.0f;
        for (int c = 0; c < C; c++) {
            num += (float)x[c] * (float)sm[c];
            den += (float)sm[c] * (float)sm[c];
        }
        if (den == 0) return 0;
        if (num <= den && num >= k->tau * den) {
            float a = num / den;
            float d2 = 0.0f;
            for (int c = 0; c < C; c++) {
                float dd = a * (float)sm[c] - (float)x[c];
                d2 += dd * dd;
            }
            if (d2 < k->Tb * a * a) {
                Ps++;
                if (Ps >= k->knn) return 2;
            }
        }
    }
    return 0;
}

ORC_API void orc_knn_apply(orc_knn *k, const uint8_t *frame, uint8_t *mask)
{
    int C = k->C, nd = C + 1;
    size_t npx = (size_t)k->H * k->W;
    int u[3];
    k->nframes++;
    orc_knn_periods(k->nframes, k->history, u);
    int refresh[3];
    for (int t = 0; t < 3; t++) { /* order S, M, L */
        refresh[t] = (k->counter[t] == u[t] - 1);
        if (refresh[t]) orc_rng_fill_u8(&k->rng, k->fresh[t], (long)npx, u[t]);
    }
    for (size_t p = 0; p < npx; p++) {
        uint8_t *mdl = k->model + p * 3 * KNN_N * nd;
        const uint8_t *x = frame + p * C;
        uint8_t include;
        int r = knn_classify(k, x, mdl, &include);
        int oS = nd * k->idx[0][p];
        int oM = nd * (k->idx[1][p] + KNN_N);
        int oL = nd * (k->idx[2][p] + 2 * KNN_N);
        if (k->nxt[2][p] == k->counter[2]) {
            memcpy(mdl + oL, mdl + oM, nd);
            k->idx[2][p] = (uint8_t)(k->idx[2][p] >= KNN_N - 1 ? 0 : k->idx[2][p] + 1);
        }
        if (refresh[2]) k->nxt[2][p] = k->fresh[2][p];
        if (k->nxt[1][p] == k->counter[1]) {
            memcpy(mdl + oM, mdl + oS, nd);
            k->idx[1][p] = (uint8_t)(k->idx[1][p] >= KNN_N - 1 ? 0 : k->idx[1][p] + 1);
        }
        if (refresh[1]) k->nxt[1][p] = k->fresh[1][p];
        if (k->nxt[0][p] == k->counter[0]) {
            memcpy(mdl + oS, x, C);
            mdl[oS + C] = include;
            k->idx[0][p] = (uint8_t)(k->idx[0][p] >= KNN_N - 1 ? 0 : k->idx[0][p] + 1);
        }
        if (refresh[0]) k->nxt[0][p] = k->fresh[0][p];
        mask[p] = r == 1 ? 0 : r == 2 ? k->shadowVal : 255;
    }
    for (int t = 0; t < 3; t++) {
        k->counter[t]++;
        if (k->counter[t] >= u[t]) k->counter[t] = 0;
    }
}

/* ------------------------------------------------------------------------- */
/* Grey-level erode / dilate with an arbitrary structuring element.
 * Reference call site: pipeline.py:473-475 (d['fun'](mask, d['kernel']), fun in
 * {cv.erode, cv.dilate}).  anchor (kh/2, kw/2), out-of-image pixels ignored. */
ORC_API void orc_morph(const uint8_t *src, int H, int W, const uint8_t *se, int kh, int kw,
                       int dilate, uint8_t *dst)
{
    int ay = kh / 2, ax = kw / 2;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int v = dilate ? 0 : 255;
            for (int i = 0; i < kh; i++) {
                int yy = y + i - ay;
                if (yy < 0 || yy >= H) continue;
                for (int j = 0; j < kw; j++) {
                    if (!se[i * kw + j]) continue;
                    int xx = x + j - ax;
                    if (xx < 0 || xx >= W) continue;
                    int s = src[(size_t)yy * W + xx];
                    if (dilate ? s > v : s < v) v = s;
                }
            }
            dst[(size_t)y * W + x] = (uint8_t)v;
        }
}

/* ------------------------------------------------------------------------- */
/* 8-connected component labelling with canonical labels (1 + raster index of the
 * component's first pixel; 0 = background) and per-component statistics.  This is the
 * device-side half of blob creation (pipeline.py:555-570); contours are rebuilt from it. */
static int32_t uf_find(int32_t *parent, int32_t a)
{
    while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; }
    return a;
}
static void uf_union(int32_t *parent, int32_t a, int32_t b)
{
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a < b) parent[b] = a; else parent[a] = b;
}

ORC_API void orc_ccl8(const uint8_t *mask, int H, int W, int32_t *labels)
{
    size_t npx = (size_t)H * W;
    int32_t *parent = (int32_t *)malloc(npx * sizeof(int32_t));
    for (size_t p = 0; p < npx; p++) parent[p] = (int32_t)p;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int32_t p = y * W + x;
            if (!mask[p]) continue;
            if (x > 0 && mask[p - 1]) uf_union(parent, p, p - 1);
            if (y > 0) {
                if (mask[p - W]) uf_union(parent, p, p - W);
                if (x > 0 && mask[p - W - 1]) uf_union(parent, p, p - W - 1);
                if (x < W - 1 && mask[p - W + 1]) uf_union(parent, p, p - W + 1);
            }
        }
    for (size_t p = 0; p < npx; p++) labels[p] = mask[p] ? uf_find(parent, (int32_t)p) + 1 : 0;
    free(parent);
}
