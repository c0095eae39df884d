/*
 * frontend_oracle.c -- CPU restatement of the stereo front-end hot path of
 * famoreno/srba-stereo-slam.  TEST INFRASTRUCTURE ONLY: nothing under
 * srba-stereo-slam_b200/ may call, link or import this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY STATUS: the arithmetic of this path lives in un-vendored
 * dependencies of the reference (OpenCV >=2.4.8,<3.0 `cv::ORB`/`BFMatcher`, stereo-vo
 * (unpinned), DBoW2 1.0.2).  None of them is present or buildable here, so every stage is
 * restated from the published algorithm and pinned as follows:
 *   - resize / FAST+NMS / Harris / IC angle / fastAtan2 / rBRIEF sampling / BFMatcher:
 *     PINNED bit-exactly against cv2 4.13 building blocks (tests/golden/ *.npz made by
 *     oracle/tools/make_golden.py; the algorithms are unchanged between 2.4 and 4.x).
 *   - 7x7 sigma=2 blur in OpenCV-2.4 8-bit fixed-point arithmetic: "parity unpinned"
 *     (4.x changed it); cross-checked only against an integer sepFilter2D.
 *   - row-band L/R matcher (stereo-vo smDescRbR): "parity unpinned"; conventions in DESIGN.md.
 *   - DBoW2 transform/score: "parity unpinned" for the traversal rule, pinned for the
 *     vocabulary content (reference demo/voc.yml.gz).
 *
 * Reference call sites this follows (relative to /root/reference/src):
 *   CSRBAStereoSLAMEstimator.cpp:112,267   processNewImagePair  -> orc_orb_extract x2 + orc_match_rowband
 *   CSRBAStereoSLAMEstimator.cpp:1135-1142 forced ORB / DescRbR / DescBF, FAST th, nlevels, nfeats
 *   CSRBAStereoSLAMEstimator.cpp:1447-1452 BFMatcher(NORM_HAMMING,false).match -> orc_match_bf
 *   CBoWManager.h:71-76,83-88,95-109       repack / add / query -> orc_bow_*
 *   srba-stereo-slam_utils.cpp:49-59,75    field lists of KeyPoint / DMatch equality
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (no FMA contraction: OpenCV's generic x86 build
 * evaluates the Harris / atan / rotation float expressions unfused).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } orc_keypoint; /* cv::KeyPoint, 28 B */
typedef struct { int32_t queryIdx, trainIdx, imgIdx; float distance; } orc_dmatch;             /* cv::DMatch, 16 B  */

static const int8_t BRIEF_PATTERN[256 * 4] = {
#include "brief_pattern_31.inc"
};

static inline int cv_round(double v) { return (int)lrint(v); } /* round-half-even, like cvRound (SSE2 cvtsd2si) */
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(double v) { int i = (int)v; return i + (i < v); }

/* ------------------------------------------------------------------------------------------
 * Pyramid geometry and per-level quotas (OpenCV 2.4 orb.cpp, ORB::operator()).
 * scaleFactor is stored as double(1.2f); getScale returns float.
 * ---------------------------------------------------------------------------------------- */
float orc_level_scale(int level, float scale_factor)
{
    return (float)pow((double)scale_factor, (double)level);
}

void orc_level_sizes(int W, int H, int nlevels, float scale_factor, int32_t* lw, int32_t* lh)
{
    for (int l = 0; l < nlevels; ++l) {
        float scale = 1 / orc_level_scale(l, scale_factor);
        lw[l] = cv_round((double)((float)W * scale));
        lh[l] = cv_round((double)((float)H * scale));
    }
}

/* OpenCV >= 3.x writes Size(cvRound(cols/scale), cvRound(rows/scale)) -- a float division instead of a multiplication
 * by the reciprocal; same sizes for every configuration in BASELINE.json, kept separate for exactness. */
void orc_level_sizes_cv4(int W, int H, int nlevels, float scale_factor, int32_t* lw, int32_t* lh)
{
    for (int l = 0; l < nlevels; ++l) {
        float scale = orc_level_scale(l, scale_factor);
        lw[l] = cv_round((double)((float)W / scale));
        lh[l] = cv_round((double)((float)H / scale));
    }
}

void orc_level_quotas(int nfeatures, int nlevels, float scale_factor, int32_t* quota)
{
    float factor = (float)(1.0 / (double)scale_factor);
    float per = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        quota[l] = cv_round(per);
        sum += quota[l];
        per *= factor;
    }
    quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
}

/* ------------------------------------------------------------------------------------------
 * 8-bit INTER_LINEAR resize (OpenCV imgproc resize(): 11-bit coefficients, the special
 * 32s->8u vertical pass).  Pinned bit-exactly against cv2.resize (IPP off).
 * ---------------------------------------------------------------------------------------- */
void orc_resize_coeffs(int ssize, int dsize, int32_t* ofs, int16_t* coef /* 2*dsize */, int is_x)
{
    double inv_scale = (double)dsize / ssize;
    double scale = 1. / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor(f);
        f -= s;
        if (is_x) {   /* only the horizontal pass clamps offset AND weight; rows are clipped later */
            if (s < 0) { f = 0; s = 0; }
            if (s >= ssize - 1) { f = 0; s = ssize - 1; }
        }
        ofs[d] = s;
        float c0 = 1.f - f, c1 = f;
        coef[2 * d] = (int16_t)cv_round(c0 * 2048);
        coef[2 * d + 1] = (int16_t)cv_round(c1 * 2048);
    }
}

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                          uint8_t* dst, int dw, int dh, int dstride)
{
    int32_t* xofs = malloc(sizeof(int32_t) * dw);
    int16_t* ia = malloc(sizeof(int16_t) * 2 * dw);
    int32_t* yofs = malloc(sizeof(int32_t) * dh);
    int16_t* ib = malloc(sizeof(int16_t) * 2 * dh);
    orc_resize_coeffs(sw, dw, xofs, ia, 1);
    orc_resize_coeffs(sh, dh, yofs, ib, 0);
    int32_t* r0 = malloc(sizeof(int32_t) * dw);
    int32_t* r1 = malloc(sizeof(int32_t) * dw);
    for (int dy = 0; dy < dh; ++dy) {
        int sy0 = yofs[dy], sy1 = yofs[dy] + 1;           /* y: rows clipped, coefficients kept */
        if (sy0 < 0) sy0 = 0; if (sy0 > sh - 1) sy0 = sh - 1;
        if (sy1 < 0) sy1 = 0; if (sy1 > sh - 1) sy1 = sh - 1;
        const uint8_t* S0 = src + (size_t)sy0 * sstride;
        const uint8_t* S1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sx;
            r0[dx] = S0[sx] * ia[2 * dx] + S0[sx1] * ia[2 * dx + 1];
            r1[dx] = S1[sx] * ia[2 * dx] + S1[sx1] * ia[2 * dx + 1];
        }
        int b0 = ib[2 * dy], b1 = ib[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; ++dx)
            D[dx] = (uint8_t)((((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2);
    }
    free(xofs); free(ia); free(yofs); free(ib); free(r0); free(r1);
}

/* 8-bit INTER_LINEAR_EXACT (pyramid_mode 1): the bit-exact resize OpenCV >= 3.4 uses for ORB's pyramid.
 * Coefficients are 8-bit fractions round((f - floor f) * 256) from f = (1/(dst/src)) * (d + 0.5) - 0.5 in double;
 * horizontal pass h = c0*p0 + c1*p1 (8.8 fixed point, exact), vertical pass (b0*h0 + b1*h1 + 2^15) >> 16;
 * outside the source the nearest edge sample is taken with full weight.  Pinned bit-exactly against
 * cv2.resize(INTER_LINEAR_EXACT). */
static void exact_coeffs(int ssize, int dsize, int32_t* ofs, int32_t* c1)
{
    double inv = (double)dsize / ssize, scale = 1.0 / inv;
    for (int v = 0; v < dsize; ++v) {
        double f = scale * ((double)v + 0.5) - 0.5;
        int i = cv_floor(f);
        ofs[v] = 0; c1[v] = 0;
        if (i >= 0 && ssize > 1) {
            if (i < ssize - 1) { ofs[v] = i; c1[v] = cv_round((f - i) * 256); }
            else ofs[v] = ssize - 1;
        }
    }
}

void orc_resize_linear_exact_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride)
{
    int32_t* xo = malloc(sizeof(int32_t) * dw); int32_t* xa = malloc(sizeof(int32_t) * dw);
    int32_t* yo = malloc(sizeof(int32_t) * dh); int32_t* ya = malloc(sizeof(int32_t) * dh);
    exact_coeffs(sw, dw, xo, xa); exact_coeffs(sh, dh, yo, ya);
    for (int dy = 0; dy < dh; ++dy) {
        int y0 = yo[dy], y1 = y0 + 1 < sh ? y0 + 1 : sh - 1, b1 = ya[dy], b0 = 256 - b1;
        const uint8_t* S0 = src + (size_t)y0 * sstride; const uint8_t* S1 = src + (size_t)y1 * sstride;
        for (int dx = 0; dx < dw; ++dx) {
            int x0 = xo[dx], x1 = x0 + 1 < sw ? x0 + 1 : sw - 1, a1 = xa[dx], a0 = 256 - a1;
            int h0 = S0[x0] * a0 + S0[x1] * a1, h1 = S1[x0] * a0 + S1[x1] * a1;
            dst[(size_t)dy * dstride + dx] = (uint8_t)((h0 * b0 + h1 * b1 + (1 << 15)) >> 16);
        }
    }
    free(xo); free(xa); free(yo); free(ya);
}

/* ------------------------------------------------------------------------------------------
 * FAST-9/16 with score and 3x3 strict non-max suppression (cv::FastFeatureDetector(th,true)).
 * s(p) = max over the 16 arcs of 9 contiguous ring pixels of min |I(ring)-I(p)| taken with one
 * sign; p is a corner iff s > th; score = s-1; scores are 0 outside the 3-px margin; a corner
 * is kept iff its score is strictly greater than all 8 neighbours'.  Output is row-major.
 * ---------------------------------------------------------------------------------------- */
static const int RING_DX[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
static const int RING_DY[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };

static int fast_arc_strength(const uint8_t* p, int stride)
{
    int d[32], best = 0;
    int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = d[k + 16] = v - p[RING_DY[k] * stride + RING_DX[k]];
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) {
            int e = d[k + j];
            if (e < mn) mn = e;
            if (e > mx) mx = e;
        }
        if (mn > best) best = mn;       /* centre brighter than the whole arc */
        if (-mx > best) best = -mx;     /* centre darker than the whole arc   */
    }
    return best;
}

/* dense score map: 0 = not a corner, else score (>= th).  The opposite-pair test only skips pixels that cannot
 * be corners (every 9-arc contains one pixel of each opposite pair); it does not change the result. */
void orc_fast_score_map(const uint8_t* img, int w, int h, int stride, int th, uint8_t* score /* w*h */)
{
    memset(score, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t* row = img + (size_t)y * stride;
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t* p = row + x;
            int v = p[0], lo = v - th, hi = v + th;
            int p0 = p[3 * stride], p8 = p[-3 * stride], p4 = p[3], p12 = p[-3];
            int br = (p0 > hi || p8 > hi) && (p4 > hi || p12 > hi);
            int dk = (p0 < lo || p8 < lo) && (p4 < lo || p12 < lo);
            if (!(br || dk)) continue;
            int s = fast_arc_strength(p, stride);
            if (s > th) score[(size_t)y * w + x] = (uint8_t)(s - 1);
        }
    }
}

/* returns the number of keypoints; xy/sc may be NULL to count only. */
int orc_fast9_nms(const uint8_t* img, int w, int h, int stride, int th,
                  int32_t* xs, int32_t* ys, int32_t* sc, int cap)
{
    uint8_t* score = malloc((size_t)w * h);
    orc_fast_score_map(img, w, h, stride, th, score);
    int n = 0;
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = score[(size_t)y * w + x];
            if (!s) continue;
            const uint8_t* q = score + (size_t)y * w + x;
            if (s > q[-1] && s > q[1] && s > q[-w - 1] && s > q[-w] && s > q[-w + 1] &&
                s > q[w - 1] && s > q[w] && s > q[w + 1]) {
                if (n < cap && xs) { xs[n] = x; ys[n] = y; sc[n] = s; }
                ++n;
            }
        }
    free(score);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Harris response on a 7x7 block (orb.cpp HarrisResponses, blockSize 7, k 0.04).
 * ---------------------------------------------------------------------------------------- */
float orc_harris(const uint8_t* img, int stride, int x, int y)
{
    const int bs = 7, r = bs / 2;
    float scale = (1 << 2) * bs * 255.0f;
    scale = 1.0f / scale;
    float scale_sq_sq = scale * scale * scale * scale;
    int a = 0, b = 0, c = 0;
    for (int i = 0; i < bs; ++i)
        for (int j = 0; j < bs; ++j) {
            const uint8_t* p = img + (size_t)(y - r + i) * stride + (x - r + j);
            int Ix = (p[1] - p[-1]) * 2 + (p[-stride + 1] - p[-stride - 1]) + (p[stride + 1] - p[stride - 1]);
            int Iy = (p[stride] - p[-stride]) * 2 + (p[stride - 1] - p[-stride - 1]) + (p[stride + 1] - p[-stride + 1]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
    float fa = (float)a, fb = (float)b, fc = (float)c;
    float t1 = fa * fb;
    float t2 = fc * fc;
    float t3 = t1 - t2;
    float s = fa + fb;
    float t4 = 0.04f * s;
    float t5 = t4 * s;
    return (t3 - t5) * scale_sq_sq;
}

/* ------------------------------------------------------------------------------------------
 * Intensity-centroid orientation (orb.cpp IC_Angle, half patch 15) + cv::fastAtan2 (degrees).
 * ---------------------------------------------------------------------------------------- */
void orc_umax(int32_t* umax /* 16 */)
{
    const int half = 15;
    int v, v0, vmax = cv_floor(half * sqrt(2.f) / 2 + 1);
    int vmin = cv_ceil(half * sqrt(2.f) / 2);
    int u[17] = { 0 };
    for (v = 0; v <= vmax; ++v) u[v] = cv_round(sqrt((double)half * half - v * v));
    for (v = half, v0 = 0; v >= vmin; --v) {
        while (u[v0] == u[v0 + 1]) ++v0;
        u[v] = v0;
        ++v0;
    }
    for (v = 0; v <= half; ++v) umax[v] = u[v];
}

float orc_fast_atan2(float y, float x)
{
    static const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    static const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    static const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    static const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

void orc_ic_moments(const uint8_t* img, int stride, int x, int y, int32_t* m10, int32_t* m01)
{
    int32_t umax[16];
    orc_umax(umax);
    const uint8_t* c = img + (size_t)y * stride + x;
    int a = 0, b = 0;
    for (int u = -15; u <= 15; ++u) a += u * c[u];
    for (int v = 1; v <= 15; ++v) {
        int vs = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int p = c[u + v * stride], m = c[u - v * stride];
            vs += p - m;
            a += u * (p + m);
        }
        b += v * vs;
    }
    *m10 = a; *m01 = b;
}

float orc_ic_angle(const uint8_t* img, int stride, int x, int y)
{
    int32_t m10, m01;
    orc_ic_moments(img, stride, x, y, &m10, &m01);
    return orc_fast_atan2((float)m01, (float)m10);
}

/* ------------------------------------------------------------------------------------------
 * 7x7 sigma=2 Gaussian blur, OpenCV-2.4 8-bit path: taps round(256*g) = 18 34 49 55 49 34 18,
 * int32 row pass, column pass (sum + 2^15) >> 16 saturated; BORDER_REFLECT_101.
 * ---------------------------------------------------------------------------------------- */
static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

void orc_gauss7_taps(int32_t* taps /* 7 */)
{
    /* getGaussianKernel(7, 2, CV_32F) then convertTo(CV_32S, 256) */
    const int n = 7; const double sigma = 2.0;
    double scale2X = -0.5 / (sigma * sigma), sum = 0;
    float k[7];
    for (int i = 0; i < n; ++i) {
        double x = i - (n - 1) * 0.5;
        double t = exp(scale2X * x * x);
        k[i] = (float)t; sum += k[i];
    }
    sum = 1. / sum;
    for (int i = 0; i < n; ++i) { k[i] = (float)(k[i] * sum); taps[i] = cv_round((double)k[i] * 256); }
}

void orc_gauss7_blur_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride)
{
    int32_t t[7];
    orc_gauss7_taps(t);
    int32_t* rows = malloc(sizeof(int32_t) * (size_t)w * h);
    int* xi = malloc(sizeof(int) * (w + 6));        /* reflect-101 column / row index tables */
    int* yi = malloc(sizeof(int) * (h + 6));
    for (int x = -3; x < w + 3; ++x) xi[x + 3] = reflect101(x, w);
    for (int y = -3; y < h + 3; ++y) yi[y + 3] = reflect101(y, h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* r = src + (size_t)y * sstride;
        int32_t* o = rows + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            const int* c = xi + x;
            o[x] = t[0] * r[c[0]] + t[1] * r[c[1]] + t[2] * r[c[2]] + t[3] * r[c[3]] + t[4] * r[c[4]] + t[5] * r[c[5]] + t[6] * r[c[6]];
        }
    }
    for (int y = 0; y < h; ++y) {
        const int32_t* r[7];
        for (int k = 0; k < 7; ++k) r[k] = rows + (size_t)yi[y + k] * w;
        uint8_t* o = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            int s = t[0] * r[0][x] + t[1] * r[1][x] + t[2] * r[2][x] + t[3] * r[3][x] + t[4] * r[4][x] + t[5] * r[5][x] + t[6] * r[6][x];
            s = (s + (1 << 15)) >> 16;
            o[x] = (uint8_t)(s > 255 ? 255 : s);
        }
    }
    free(rows); free(xi); free(yi);
}

/* Alternative blur arithmetic (blur_mode 1): what cv::sepFilter2D does since OpenCV 3.4/4.x when
 * the float Gaussian taps are not exact multiples of 1/256 -- float32 taps, float row pass
 * (k0*S0 + k1*S1 + ... left to right), float symmetric column pass (kc*Sc, then += k[j]*(S[+j]+S[-j])),
 * round-half-even, saturate.  This is the blur inside cv2 4.13's ORB, so mode 1 makes the whole
 * nlevels=1 extraction bit-exact against cv2.ORB (PINNED, tests/golden/orb_nl1_*.npz). */
void orc_gauss7_taps_f32(float* k /* 7 */)
{
    const int n = 7; const double sigma = 2.0;
    double scale2X = -0.5 / (sigma * sigma), sum = 0;
    for (int i = 0; i < n; ++i) {
        double x = i - (n - 1) * 0.5;
        k[i] = (float)exp(scale2X * x * x); sum += k[i];
    }
    sum = 1. / sum;
    for (int i = 0; i < n; ++i) k[i] = (float)(k[i] * sum);
}

void orc_gauss7_blur_f32path_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride)
{
    float k[7];
    orc_gauss7_taps_f32(k);
    float* rows = malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float s = k[0] * (float)src[(size_t)y * sstride + reflect101(x - 3, w)];
            for (int i = 1; i < 7; ++i) s += k[i] * (float)src[(size_t)y * sstride + reflect101(x - 3 + i, w)];
            rows[(size_t)y * w + x] = s;
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float s = k[3] * rows[(size_t)y * w + x];
            for (int j = 1; j <= 3; ++j)
                s += k[3 + j] * (rows[(size_t)reflect101(y + j, h) * w + x] + rows[(size_t)reflect101(y - j, h) * w + x]);
            int v = cv_round((double)s);
            dst[(size_t)y * dstride + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    free(rows);
}

/* ------------------------------------------------------------------------------------------
 * Steered BRIEF (orb.cpp computeOrbDescriptor, WTA_K=2): bit i = I(rot a_i) < I(rot b_i) on the
 * blurred level, LSB-first in byte i/8.  angle in degrees.
 * ---------------------------------------------------------------------------------------- */
void orc_brief(const uint8_t* blurred, int stride, int x, int y, float angle_deg, uint8_t* desc /* 32 */)
{
    float angle = angle_deg;
    angle *= (float)(M_PI / 180.f);
    float a = (float)cos(angle), b = (float)sin(angle);
    const uint8_t* c = blurred + (size_t)y * stride + x;
    for (int i = 0; i < 32; ++i) {
        int byte = 0;
        for (int j = 0; j < 8; ++j) {
            const int8_t* p = BRIEF_PATTERN + 4 * (8 * i + j);
            float x0 = p[0] * a - p[1] * b, y0 = p[0] * b + p[1] * a;
            float x1 = p[2] * a - p[3] * b, y1 = p[2] * b + p[3] * a;
            int t0 = c[cv_round(y0) * stride + cv_round(x0)];
            int t1 = c[cv_round(y1) * stride + cv_round(x1)];
            byte |= (t0 < t1) << j;
        }
        desc[i] = (uint8_t)byte;
    }
}

const int8_t* orc_brief_pattern(void) { return BRIEF_PATTERN; }

/* ------------------------------------------------------------------------------------------
 * retainBest with ties (KeyPointsFilter::retainBest as documented: keep the n best plus every
 * point whose response equals the n-th best).  Returns the boundary response in *boundary;
 * keep[i] = response[i] >= boundary.  n >= count keeps all.
 * ---------------------------------------------------------------------------------------- */
static int cmp_float_desc(const void* a, const void* b)
{
    float x = *(const float*)a, y = *(const float*)b;
    return x > y ? -1 : x < y ? 1 : 0;
}

int orc_retain_best(const float* resp, int count, int n, uint8_t* keep)
{
    if (n >= count) { memset(keep, 1, count); return count; }
    if (n <= 0) { memset(keep, 0, count); return 0; }
    float* s = malloc(sizeof(float) * count);
    memcpy(s, resp, sizeof(float) * count);
    qsort(s, count, sizeof(float), cmp_float_desc);
    float boundary = s[n - 1];
    free(s);
    int kept = 0;
    for (int i = 0; i < count; ++i) { keep[i] = resp[i] >= boundary; kept += keep[i]; }
    return kept;
}

/* ------------------------------------------------------------------------------------------
 * Whole-image ORB (cv::ORB::operator() of OpenCV 2.4 with a run-time FAST threshold).
 * Canonical keypoint order: level-major, then (y, x) ascending in level coordinates -- the
 * reference's own order is an std::nth_element artefact.
 * Outputs: kps[cap], desc[cap*32]; optional pyramid/blurred level dumps (concatenated tightly,
 * level after level) for pixel-exact tests.  Returns the keypoint count (may exceed cap: then
 * only the first cap are stored).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t nfeatures, nlevels, fast_threshold, edge_threshold;
    float scale_factor;
    int32_t blur_mode;   /* 0: OpenCV-2.4 8-bit fixed point (reference parity); 1: float path (cv >= 3.4) */
    int32_t pyramid_mode;/* 0: OpenCV-2.4 INTER_LINEAR (reference parity); 1: INTER_LINEAR_EXACT + cv4 level sizes (cv >= 3.4) */
} orc_orb_params;

int orc_orb_extract(const uint8_t* img, int W, int H, int stride, const orc_orb_params* prm,
                    orc_keypoint* kps, uint8_t* desc, int cap,
                    uint8_t* pyr_out /* may be NULL */, uint8_t* blur_out /* may be NULL */,
                    int32_t* level_counts /* nlevels, may be NULL */)
{
    int L = prm->nlevels;
    int32_t lw[32], lh[32], quota[32];
    if (prm->pyramid_mode == 1) orc_level_sizes_cv4(W, H, L, prm->scale_factor, lw, lh);
    else orc_level_sizes(W, H, L, prm->scale_factor, lw, lh);
    orc_level_quotas(prm->nfeatures, L, prm->scale_factor, quota);
    uint8_t* prev = NULL; int n = 0;
    size_t out_off = 0;
    for (int l = 0; l < L; ++l) {
        int w = lw[l], h = lh[l];
        uint8_t* cur = malloc((size_t)w * h);
        if (l == 0) for (int y = 0; y < h; ++y) memcpy(cur + (size_t)y * w, img + (size_t)y * stride, w);
        else if (prm->pyramid_mode == 1) orc_resize_linear_exact_u8(prev, lw[l - 1], lh[l - 1], lw[l - 1], cur, w, h, w);
        else orc_resize_linear_u8(prev, lw[l - 1], lh[l - 1], lw[l - 1], cur, w, h, w);
        if (pyr_out) memcpy(pyr_out + out_off, cur, (size_t)w * h);
        int level_n = 0;
        int e = prm->edge_threshold;
        if (w > 2 * e && h > 2 * e && w >= 7 && h >= 7) {
            int cap_raw = ((w + 1) / 2) * ((h + 1) / 2);    /* NMS: no two 8-adjacent maxima */
            int32_t* xs = malloc(sizeof(int32_t) * (cap_raw + 1));
            int32_t* ys = malloc(sizeof(int32_t) * (cap_raw + 1));
            int32_t* sc = malloc(sizeof(int32_t) * (cap_raw + 1));
            cap_raw = orc_fast9_nms(cur, w, h, w, prm->fast_threshold, xs, ys, sc, cap_raw);
            /* runByImageBorder(edgeThreshold) */
            int m = 0;
            for (int i = 0; i < cap_raw; ++i)
                if (xs[i] >= e && xs[i] < w - e && ys[i] >= e && ys[i] < h - e) { xs[m] = xs[i]; ys[m] = ys[i]; sc[m] = sc[i]; ++m; }
            float* resp = malloc(sizeof(float) * (m + 1));
            uint8_t* keep = malloc(m + 1);
            for (int i = 0; i < m; ++i) resp[i] = (float)sc[i];
            orc_retain_best(resp, m, 2 * quota[l], keep);
            int m2 = 0;
            for (int i = 0; i < m; ++i) if (keep[i]) { xs[m2] = xs[i]; ys[m2] = ys[i]; ++m2; }
            for (int i = 0; i < m2; ++i) resp[i] = orc_harris(cur, w, xs[i], ys[i]);
            orc_retain_best(resp, m2, quota[l], keep);
            float sf = orc_level_scale(l, prm->scale_factor);
            /* descriptors are sampled from the blurred level (blur happens after detection) */
            uint8_t* bl = malloc((size_t)w * h);
            if (prm->blur_mode == 1) orc_gauss7_blur_f32path_u8(cur, w, h, w, bl, w);
            else orc_gauss7_blur_u8(cur, w, h, w, bl, w);
            if (blur_out) memcpy(blur_out + out_off, bl, (size_t)w * h);
            for (int i = 0; i < m2; ++i) {
                if (!keep[i]) continue;
                if (n < cap) {
                    orc_keypoint* k = kps + n;
                    float ang = orc_ic_angle(cur, w, xs[i], ys[i]);
                    k->x = (float)xs[i]; k->y = (float)ys[i];
                    if (l != 0) { k->x *= sf; k->y *= sf; }
                    k->size = 31 * sf; k->angle = ang; k->response = resp[i];
                    k->octave = l; k->class_id = -1;
                    orc_brief(bl, w, xs[i], ys[i], ang, desc + (size_t)n * 32);
                }
                ++n; ++level_n;
            }
            free(bl); free(resp); free(keep); free(xs); free(ys); free(sc);
        } else if (blur_out) {
            if (prm->blur_mode == 1) orc_gauss7_blur_f32path_u8(cur, w, h, w, blur_out + out_off, w);
            else orc_gauss7_blur_u8(cur, w, h, w, blur_out + out_off, w);
        }
        if (level_counts) level_counts[l] = level_n;
        out_off += (size_t)w * h;
        free(prev); prev = cur;
    }
    free(prev);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Brute-force Hamming match: cv::BFMatcher(NORM_HAMMING,false).match (EST.cpp:1447-1452).
 * Lowest train index wins ties; distance is a float; empty train set -> no matches.
 * ---------------------------------------------------------------------------------------- */
static inline int hamming256(const uint8_t* a, const uint8_t* b)
{
    const uint64_t* p = (const uint64_t*)a; const uint64_t* q = (const uint64_t*)b;
    uint64_t x0, x1, x2, x3, y0, y1, y2, y3;
    memcpy(&x0, p, 8); memcpy(&x1, p + 1, 8); memcpy(&x2, p + 2, 8); memcpy(&x3, p + 3, 8);
    memcpy(&y0, q, 8); memcpy(&y1, q + 1, 8); memcpy(&y2, q + 2, 8); memcpy(&y3, q + 3, 8);
    return __builtin_popcountll(x0 ^ y0) + __builtin_popcountll(x1 ^ y1) +
           __builtin_popcountll(x2 ^ y2) + __builtin_popcountll(x3 ^ y3);
}

int orc_match_bf(const uint8_t* Q, int nq, const uint8_t* T, int nt, orc_dmatch* out)
{
    if (nt <= 0) return 0;
    for (int q = 0; q < nq; ++q) {
        int best = 1 << 30, bi = -1;
        for (int t = 0; t < nt; ++t) {
            int d = hamming256(Q + (size_t)q * 32, T + (size_t)t * 32);
            if (d < best) { best = d; bi = t; }
        }
        out[q].queryIdx = q; out[q].trainIdx = bi; out[q].imgIdx = 0; out[q].distance = (float)best;
    }
    return nq;
}

/* ------------------------------------------------------------------------------------------
 * Row-band left/right match (stereo-vo smDescRbR; "parity unpinned", conventions in DESIGN.md):
 * candidates for left i: right j with |yl - yr| <= max_y_diff and min_disp <= xl - xr <= max_disp;
 * best = min Hamming, lowest j on ties; accepted iff dist <= max_dist.  robust_1to1: additionally
 * i must be the best (lowest i on ties) left for j among lefts whose best is j.
 * Output in left order.  Returns the number of matches.
 * ---------------------------------------------------------------------------------------- */
int orc_match_rowband(const orc_keypoint* kl, const uint8_t* dl, int nl,
                      const orc_keypoint* kr, const uint8_t* dr, int nr,
                      float max_y_diff, float min_disp, float max_disp, int max_dist, int robust_1to1,
                      orc_dmatch* out)
{
    int* best_j = malloc(sizeof(int) * (nl + 1));
    int* best_d = malloc(sizeof(int) * (nl + 1));
    for (int i = 0; i < nl; ++i) {
        int bd = 1 << 30, bj = -1;
        for (int j = 0; j < nr; ++j) {
            float dy = fabsf(kl[i].y - kr[j].y);
            float disp = kl[i].x - kr[j].x;
            if (!(dy <= max_y_diff) || !(disp >= min_disp) || !(disp <= max_disp)) continue;
            int d = hamming256(dl + (size_t)i * 32, dr + (size_t)j * 32);
            if (d < bd) { bd = d; bj = j; }
        }
        if (bj >= 0 && bd > max_dist) bj = -1;
        best_j[i] = bj; best_d[i] = bd;
    }
    if (robust_1to1) {
        int* owner = malloc(sizeof(int) * (nr + 1));
        for (int j = 0; j < nr; ++j) owner[j] = -1;
        for (int i = 0; i < nl; ++i) {
            int j = best_j[i];
            if (j < 0) continue;
            if (owner[j] < 0 || best_d[i] < best_d[owner[j]]) owner[j] = i;
        }
        for (int i = 0; i < nl; ++i) if (best_j[i] >= 0 && owner[best_j[i]] != i) best_j[i] = -1;
        free(owner);
    }
    int m = 0;
    for (int i = 0; i < nl; ++i) {
        if (best_j[i] < 0) continue;
        out[m].queryIdx = i; out[m].trainIdx = best_j[i]; out[m].imgIdx = 0; out[m].distance = (float)best_d[i];
        ++m;
    }
    free(best_j); free(best_d);
    return m;
}

/* ------------------------------------------------------------------------------------------
 * DBoW2 1.0.2 vocabulary transform and L1 scoring (CBoWManager.h:71-88).
 * Tree arrays are indexed by DBoW2 node id (0 = root); children listed in file order.
 * desc bytes are in OpenCV order: DBoW2 bit i <-> byte i/8, bit i%8 (CBoWManager.h:102-106).
 * ---------------------------------------------------------------------------------------- */
void orc_bow_transform(const int32_t* child_ofs /* nnodes+1 */, const int32_t* child_idx,
                       const uint8_t* node_desc /* nnodes*32 */, const double* node_weight,
                       const int32_t* node_word /* -1 for inner */,
                       const uint8_t* desc, int n, int32_t* word, double* weight)
{
    for (int i = 0; i < n; ++i) {
        int cur = 0;
        do {
            int best = 1 << 30, bc = -1;
            for (int c = child_ofs[cur]; c < child_ofs[cur + 1]; ++c) {
                int id = child_idx[c];
                int d = hamming256(desc + (size_t)i * 32, node_desc + (size_t)id * 32);
                if (d < best) { best = d; bc = id; }
            }
            cur = bc;
        } while (child_ofs[cur] != child_ofs[cur + 1]);
        word[i] = node_word[cur];
        weight[i] = node_weight[cur];
    }
}

/* BowVector: v[word] += w for w > 0 (TF-IDF), then L1-normalised.  Output sorted by word id.
 * Returns the number of distinct words. */
int orc_bow_vector(const int32_t* word, const double* weight, int n, int nwords,
                   int32_t* out_word, double* out_val)
{
    double* acc = calloc(nwords, sizeof(double));
    uint8_t* seen = calloc(nwords, 1);
    for (int i = 0; i < n; ++i)
        if (weight[i] > 0) { acc[word[i]] += weight[i]; seen[word[i]] = 1; }
    int m = 0; double norm = 0;
    for (int w = 0; w < nwords; ++w) if (seen[w]) { out_word[m] = w; out_val[m] = acc[w]; norm += fabs(acc[w]); ++m; }
    if (norm > 0) for (int i = 0; i < m; ++i) out_val[i] /= norm;
    free(acc); free(seen);
    return m;
}

/* L1 score of DBoW2's queryL1 between two sorted sparse vectors, accumulated in ascending word
 * order: sum over common words of |q-d|-|q|-|d|; *common = 1 iff any word is shared.
 * DBoW2 reports -sum/2. */
double orc_bow_l1_raw(const int32_t* qw, const double* qv, int nq,
                      const int32_t* dw, const double* dv, int nd, int* common)
{
    int i = 0, j = 0; double s = 0; *common = 0;
    while (i < nq && j < nd) {
        if (qw[i] == dw[j]) { s += fabs(qv[i] - dv[j]) - fabs(qv[i]) - fabs(dv[j]); *common = 1; ++i; ++j; }
        else if (qw[i] < dw[j]) ++i; else ++j;
    }
    return s;
}
