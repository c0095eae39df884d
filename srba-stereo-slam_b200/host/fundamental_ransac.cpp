// fundamental_ransac.cpp -- see fundamental_ransac.h.  Host code by design: N is 15 .. a few hundred correspondences per
// keyframe pair, the sampling loop is sequential through one RNG stream, and the inlier mask must be reproducible.
#include "fundamental_ransac.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

// ---- null space of the 7x9 epipolar system ---------------------------------------------------------------------------
// Householder QR of A^T (9x7): A^T = Q R, so the last two columns of Q span null(A).  Backward stable on raw pixel
// coordinates (entries up to ~1e6 next to 1), which a normal-equations eigen-solve is not.  Sign convention: each
// reflector maps its column onto -sign(x0)*|x| e0.  The test-suite's independent restatement follows the same steps.
static void nullspace_7x9(const double A[7][9], double f1[9], double f2[9])
{
    double M[9][7];                                   // A^T, overwritten by R; reflectors kept in V
    for (int i = 0; i < 7; ++i) for (int j = 0; j < 9; ++j) M[j][i] = A[i][j];
    double V[7][9]; double beta[7];
    for (int k = 0; k < 7; ++k) {
        double nrm = 0; for (int i = k; i < 9; ++i) nrm += M[i][k] * M[i][k];
        nrm = std::sqrt(nrm);
        for (int i = 0; i < 9; ++i) V[k][i] = 0;
        if (nrm == 0.0) { beta[k] = 0; continue; }
        const double alpha = M[k][k] >= 0 ? -nrm : nrm;
        double vn = 0;
        for (int i = k; i < 9; ++i) { V[k][i] = M[i][k]; }
        V[k][k] -= alpha;
        for (int i = k; i < 9; ++i) vn += V[k][i] * V[k][i];
        beta[k] = vn > 0 ? 2.0 / vn : 0.0;
        for (int j = k; j < 7; ++j) {                 // M <- (I - beta v v^T) M
            double s = 0; for (int i = k; i < 9; ++i) s += V[k][i] * M[i][j];
            s *= beta[k];
            for (int i = k; i < 9; ++i) M[i][j] -= s * V[k][i];
        }
    }
    // columns 7 and 8 of Q = H0 H1 ... H6 applied to e7, e8
    double* out[2] = { f1, f2 };
    for (int c = 0; c < 2; ++c) {
        double q[9]; for (int i = 0; i < 9; ++i) q[i] = i == 7 + c ? 1.0 : 0.0;
        for (int k = 6; k >= 0; --k) {
            double s = 0; for (int i = k; i < 9; ++i) s += V[k][i] * q[i];
            s *= beta[k];
            for (int i = k; i < 9; ++i) q[i] -= s * V[k][i];
        }
        for (int i = 0; i < 9; ++i) out[c][i] = q[i];
    }
}

static inline double det3(const double* m)
{
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// Real roots of c0 x^3 + c1 x^2 + c2 x + c3 = 0 (trigonometric form for three real roots, Cardano for one), in the
// order cvSolveCubic emits them.
static int solve_cubic(const double c[4], double r[3])
{
    const double a0 = c[0], b0 = c[1], c0 = c[2], d0 = c[3];
    if (a0 == 0) {
        if (b0 == 0) { if (c0 == 0) return d0 == 0 ? -1 : 0; r[0] = -d0 / c0; return 1; }
        double d = c0 * c0 - 4 * b0 * d0;
        if (d < 0) return 0;
        d = std::sqrt(d);
        const double q1 = (-c0 + d) * 0.5, q2 = (-c0 - d) * 0.5;
        if (std::fabs(q1) > std::fabs(q2)) { r[0] = q1 / b0; r[1] = d0 / q1; } else { r[0] = q2 / b0; r[1] = d0 / q2; }
        return d > 0 ? 2 : 1;
    }
    const double a1 = b0 / a0, a2 = c0 / a0, a3 = d0 / a0;
    const double Q = (a1 * a1 - 3 * a2) * (1. / 9), R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    const double Qcubed = Q * Q * Q;
    double d = Qcubed - R * R;
    if (d >= 0) {
        if (Qcubed == 0) { r[0] = r[1] = r[2] = -a1 / 3; return 3; }     // triple root
        const double theta = std::acos(std::max(-1.0, std::min(1.0, R / std::sqrt(Qcubed))));
        const double t0 = -2 * std::sqrt(Q), t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
        r[0] = t0 * std::cos(t1) - t2;
        r[1] = t0 * std::cos(t1 + (2. * M_PI / 3)) - t2;
        r[2] = t0 * std::cos(t1 + (4. * M_PI / 3)) - t2;
        return 3;
    }
    d = std::sqrt(-d);
    double e = std::pow(d + std::fabs(R), 0.333333333333);
    if (R > 0) e = -e;
    r[0] = (e + Q / e) - a1 * (1. / 3);
    return 1;
}

int fundamental7Point(const double* m1, const double* m2, double* Fout)
{
    double A[7][9];
    for (int i = 0; i < 7; ++i) {                     // (m2, 1)^T F (m1, 1) = 0, F row-major
        const double x0 = m1[2 * i], y0 = m1[2 * i + 1], x1 = m2[2 * i], y1 = m2[2 * i + 1];
        const double row[9] = { x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, 1.0 };
        for (int j = 0; j < 9; ++j) A[i][j] = row[j];
    }
    double g1[9], g2[9];
    nullspace_7x9(A, g1, g2);
    // F(l) = l*g1 + (1-l)*g2 = g2 + l*(g1-g2);  det F(l) is a cubic in l.  Its coefficients come from four evaluations
    // (l = 0, 1, -1, 2): exact for a cubic and needs nothing but 3x3 determinants.
    double dlt[9]; for (int i = 0; i < 9; ++i) dlt[i] = g1[i] - g2[i];
    auto detAt = [&](double l) { double m[9]; for (int i = 0; i < 9; ++i) m[i] = g2[i] + l * dlt[i]; return det3(m); };
    const double p0 = detAt(0), p1 = detAt(1), pm = detAt(-1), p2 = detAt(2);
    double c[4];
    c[3] = p0;
    c[1] = 0.5 * (p1 + pm) - p0;                      // l^2
    // p(l) = c0 l^3 + c1 l^2 + c2 l + c3:  p1 - pm = 2 c0 + 2 c2;  p2 = 8 c0 + 4 c1 + 2 c2 + c3
    const double s = 0.5 * (p1 - pm);                 // c0 + c2
    c[0] = (p2 - 4 * c[1] - c[3] - 2 * s) / 6.0;
    c[2] = s - c[0];
    double roots[3];
    const int n = solve_cubic(c, roots);
    if (n < 1 || n > 3) return n < 0 ? 0 : n;
    for (int k = 0; k < n; ++k, Fout += 9) {
        double lambda = roots[k], mu = 1.0;
        const double w = dlt[8] * roots[k] + g2[8];
        if (std::fabs(w) > DBL_EPSILON) { mu = 1.0 / w; lambda *= mu; Fout[8] = 1.0; } else Fout[8] = 0.0;   // scale so that F33 = 1
        for (int i = 0; i < 8; ++i) Fout[i] = dlt[i] * lambda + g2[i] * mu;
    }
    return n;
}

// modelest.cpp, CvModelEstimator2::checkSubset with checkPartialSubsets: the newest point must not lie on a line through
// two earlier points of the sample.
static bool newest_point_ok(const double* pts, int count)
{
    if (count <= 2) return true;
    const int i = count - 1;
    for (int j = 0; j < i; ++j) {
        const double dx1 = pts[2 * j] - pts[2 * i], dy1 = pts[2 * j + 1] - pts[2 * i + 1];
        for (int k = 0; k < j; ++k) {
            const double dx2 = pts[2 * k] - pts[2 * i], dy2 = pts[2 * k + 1] - pts[2 * i + 1];
            if (std::fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (std::fabs(dx1) + std::fabs(dy1) + std::fabs(dx2) + std::fabs(dy2))) return false;
        }
    }
    return true;
}

// OpenCV >= 3 (ptsetreg.cpp RANSACPointSetRegistrator::getSubset + FMEstimatorCallback::checkSubset): all seven indices are
// drawn first (duplicates redrawn), then ONE check that the last point is not collinear with any earlier pair; a failed
// check redraws the whole sample.
static bool draw_subset_cv4(CvRngMwc& rng, const double* m1, const double* m2, int count, double* s1, double* s2, int maxAttempts)
{
    int idx[7];
    for (int iters = 0; iters < maxAttempts; ++iters) {
        for (int i = 0; i < 7; ++i) {
            int id;
            for (;;) { id = (int)(rng.next() % (uint32_t)count); int j = 0; for (; j < i; ++j) if (idx[j] == id) break; if (j == i) break; }
            idx[i] = id;
            s1[2 * i] = m1[2 * id]; s1[2 * i + 1] = m1[2 * id + 1]; s2[2 * i] = m2[2 * id]; s2[2 * i + 1] = m2[2 * id + 1];
        }
        if (newest_point_ok(s1, 7) && newest_point_ok(s2, 7)) return true;
    }
    return false;
}

static bool draw_subset(CvRngMwc& rng, const double* m1, const double* m2, int count, double* s1, double* s2, int maxAttempts)
{
    const int modelPoints = 7;
    int idx[7]; int i = 0, iters = 0;
    for (; iters < maxAttempts; ++iters) {
        for (i = 0; i < modelPoints && iters < maxAttempts;) {
            const int id = (int)(rng.next() % (uint32_t)count);
            idx[i] = id;
            int j = 0; for (; j < i; ++j) if (id == idx[j]) break;
            if (j < i) continue;
            s1[2 * i] = m1[2 * id]; s1[2 * i + 1] = m1[2 * id + 1]; s2[2 * i] = m2[2 * id]; s2[2 * i + 1] = m2[2 * id + 1];
            if (!newest_point_ok(s1, i + 1) || !newest_point_ok(s2, i + 1)) { ++iters; continue; }
            ++i;
        }
        break;
    }
    return i == modelPoints && iters < maxAttempts;
}

static int count_inliers(const double* m1, const double* m2, int count, const double* F, double thr2, uint8_t* mask)
{
    int good = 0;
    for (int i = 0; i < count; ++i) {
        const double x1 = m1[2 * i], y1 = m1[2 * i + 1], x2 = m2[2 * i], y2 = m2[2 * i + 1];
        double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5], c = F[6] * x1 + F[7] * y1 + F[8];
        const double s2 = 1. / (a * a + b * b), d2 = x2 * a + y2 * b + c;
        a = F[0] * x2 + F[3] * y2 + F[6]; b = F[1] * x2 + F[4] * y2 + F[7]; c = F[2] * x2 + F[5] * y2 + F[8];
        const double s1 = 1. / (a * a + b * b), d1 = x1 * a + y1 * b + c;
        const float err = (float)std::max(d1 * d1 * s1, d2 * d2 * s2);
        const uint8_t f = err <= thr2;
        mask[i] = f; good += f;
    }
    return good;
}

static int ransac_update_iters(double p, double ep, int model_points, int max_iters)
{
    p = std::min(std::max(p, 0.), 1.); ep = std::min(std::max(ep, 0.), 1.);
    double num = std::max(1. - p, DBL_MIN), denom = 1. - std::pow(1. - ep, model_points);
    if (denom < DBL_MIN) return 0;
    num = std::log(num); denom = std::log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)std::lrint(num / denom);
}

int findFundamentalMatRansac(const float* p1, const float* p2, int n, double param1, double param2, uint8_t* mask_out, double* F_out, int* iters_out)
{
    return findFundamentalMatRansacEx(p1, p2, n, param1, param2, mask_out, F_out, iters_out, SRBA_FM_CV24);
}

int findFundamentalMatRansacEx(const float* p1, const float* p2, int n, double param1, double param2, uint8_t* mask_out, double* F_out, int* iters_out,
                               int generation)
{
    if (mask_out) std::memset(mask_out, 0, (size_t)std::max(n, 0));
    if (iters_out) *iters_out = 0;
    if (n < 15 || !p1 || !p2 || !mask_out) return 0;          // below 15 points OpenCV 2.4 switches to LMedS; the reference never calls it there
    if (param1 <= 0) param1 = 3;
    if (param2 < DBL_EPSILON || param2 > 1 - DBL_EPSILON) param2 = 0.99;
    std::vector<double> m1(2 * (size_t)n), m2(2 * (size_t)n);
    for (int i = 0; i < 2 * n; ++i) { m1[i] = p1[i]; m2[i] = p2[i]; }
    const bool cv4 = generation == SRBA_FM_CV4;
    const int modelPoints = 7, maxIters = cv4 ? 1000 : 2000;
    const double thr2 = param1 * param1;
    CvRngMwc rng(0xffffffffffffffffull);                      // cvRNG(-1), re-seeded for every call
    std::vector<uint8_t> tmask(n), best(n, 0);
    double models[27], bestF[9] = { 0 }, s1[14], s2[14];
    int niters = maxIters, maxGood = 0, iter = 0;
    for (; iter < niters; ++iter) {
        const bool drawn = cv4 ? draw_subset_cv4(rng, m1.data(), m2.data(), n, s1, s2, 10000) : draw_subset(rng, m1.data(), m2.data(), n, s1, s2, 300);
        if (!drawn) { if (iter == 0) return 0; break; }
        const int nmodels = fundamental7Point(s1, s2, models);
        if (nmodels <= 0) continue;
        for (int k = 0; k < nmodels; ++k) {
            const int good = count_inliers(m1.data(), m2.data(), n, models + 9 * k, thr2, tmask.data());
            if (good > std::max(maxGood, modelPoints - 1)) {
                best.swap(tmask); tmask.resize(n);
                std::memcpy(bestF, models + 9 * k, sizeof(bestF));
                maxGood = good;
                niters = ransac_update_iters(param2, (double)(n - good) / n, modelPoints, niters);
            }
        }
    }
    if (iters_out) *iters_out = iter;
    if (maxGood <= 0) return 0;
    std::memcpy(mask_out, best.data(), (size_t)n);
    if (F_out) std::memcpy(F_out, bestF, sizeof(bestF));
    return maxGood;
}

extern "C" {
// ctypes shims for the CPU test-suite
int srba_find_fundamental_ransac(const float* p1, const float* p2, int n, double param1, double param2, uint8_t* mask, double* F, int* iters)
{ return findFundamentalMatRansac(p1, p2, n, param1, param2, mask, F, iters); }
int srba_find_fundamental_ransac_ex(const float* p1, const float* p2, int n, double param1, double param2, uint8_t* mask, double* F, int* iters, int generation)
{ return findFundamentalMatRansacEx(p1, p2, n, param1, param2, mask, F, iters, generation); }
int srba_fundamental_7point(const double* m1, const double* m2, double* F) { return fundamental7Point(m1, m2, F); }
uint32_t srba_cv_rng_next(uint64_t* state) { CvRngMwc r(*state); const uint32_t v = r.next(); *state = r.state; return v; }
}
