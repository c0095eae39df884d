#include "stereo_pose_solver.h"
#include <cmath>
#include <cstring>

namespace {
struct Rot { double m[9]; };
Rot rodrigues(const double w[3])
{
    const double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double a, b;
    if (th < 1e-8) { a = 1.0 - th * th / 6.0; b = 0.5 - th * th / 24.0; } else { a = std::sin(th) / th; b = (1.0 - std::cos(th)) / (th * th); }
    const double x = w[0], y = w[1], z = w[2];
    Rot R = { { 1 - b * (y * y + z * z), -a * z + b * x * y, a * y + b * x * z,
                a * z + b * x * y, 1 - b * (x * x + z * z), -a * x + b * y * z,
                -a * y + b * x * z, a * x + b * y * z, 1 - b * (x * x + y * y) } };
    return R;
}
void rot_log(const Rot& R, double w[3])
{
    double c = 0.5 * (R.m[0] + R.m[4] + R.m[8] - 1.0);
    c = c > 1 ? 1 : (c < -1 ? -1 : c);
    const double th = std::acos(c);
    const double ax = R.m[7] - R.m[5], ay = R.m[2] - R.m[6], az = R.m[3] - R.m[1];
    const double k = th < 1e-8 ? 0.5 : th / (2.0 * std::sin(th));      // frame-to-frame rotations stay far from pi
    w[0] = k * ax; w[1] = k * ay; w[2] = k * az;
}
Rot mul(const Rot& A, const Rot& B)
{
    Rot C;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
    return C;
}
// In-place Cholesky solve of the 6x6 system H x = b (H symmetric positive definite).  false if H is not SPD.
bool chol6(double H[36], double b[6])
{
    for (int j = 0; j < 6; ++j) {
        double d = H[7 * j];
        for (int k = 0; k < j; ++k) d -= H[6 * j + k] * H[6 * j + k];
        if (!(d > 1e-12)) return false;
        d = std::sqrt(d); H[7 * j] = d;
        for (int i = j + 1; i < 6; ++i) {
            double s = H[6 * i + j];
            for (int k = 0; k < j; ++k) s -= H[6 * i + k] * H[6 * j + k];
            H[6 * i + j] = s / d;
        }
    }
    for (int i = 0; i < 6; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= H[6 * i + k] * b[k]; b[i] = s / H[7 * i]; }
    for (int i = 5; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < 6; ++k) s -= H[6 * k + i] * b[k]; b[i] = s / H[7 * i]; }
    return true;
}
struct Problem {
    const double* cur; int n; TStereoPoseCamera cam; std::vector<double> P; std::vector<uint8_t> ok;   // triangulated previous points
};
// squared residual of pair i at (R, t); false if the point does not project
bool residual_at(const Problem& pb, int i, const Rot& R, const double t[3], double r[3], double Q[3])
{
    const double* p = &pb.P[3 * i];
    for (int a = 0; a < 3; ++a) Q[a] = R.m[3 * a] * p[0] + R.m[3 * a + 1] * p[1] + R.m[3 * a + 2] * p[2] + t[a];
    if (!(Q[2] > 1e-6)) return false;
    const double iz = 1.0 / Q[2];
    r[0] = pb.cam.fl * Q[0] * iz + pb.cam.cul - pb.cur[3 * i];
    r[1] = pb.cam.fl * Q[1] * iz + pb.cam.cvl - pb.cur[3 * i + 1];
    r[2] = pb.cam.fr * (Q[0] - pb.cam.baseline) * iz + pb.cam.cur - pb.cur[3 * i + 2];
    return true;
}
int run_stage(const Problem& pb, const std::vector<uint8_t>& active, int max_iters, const TStereoPoseSolverParams& prm, Rot& R, double t[3], bool& failed)
{
    int incr = 0, nit = 0; double prev_cost = 0; failed = false;
    const double k2 = prm.kernel_param * prm.kernel_param;
    for (int it = 0; it < max_iters; ++it) {
        double H[36], g[6], cost = 0;
        std::memset(H, 0, sizeof(H)); std::memset(g, 0, sizeof(g));
        for (int i = 0; i < pb.n; ++i) {
            if (!active[i]) continue;
            double r[3], Q[3];
            if (!residual_at(pb, i, R, t, r, Q)) continue;
            const double e2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
            double w = 1.0;
            if (prm.use_robust_kernel) { const double s = std::sqrt(1.0 + e2 / k2); w = 1.0 / s; cost += 2.0 * k2 * (s - 1.0); } else cost += e2;
            const double iz = 1.0 / Q[2], fl = pb.cam.fl, fr = pb.cam.fr, xb = Q[0] - pb.cam.baseline;
            // d(residual)/dQ rows, then dQ/d(dw) = -[Q]x, dQ/d(dv) = I
            const double dq[3][3] = { { fl * iz, 0, -fl * Q[0] * iz * iz }, { 0, fl * iz, -fl * Q[1] * iz * iz }, { fr * iz, 0, -fr * xb * iz * iz } };
            double J[3][6];
            for (int a = 0; a < 3; ++a) {
                const double dx = dq[a][0], dy = dq[a][1], dz = dq[a][2];
                J[a][0] = dz * Q[1] - dy * Q[2]; J[a][1] = dx * Q[2] - dz * Q[0]; J[a][2] = dy * Q[0] - dx * Q[1];
                J[a][3] = dx; J[a][4] = dy; J[a][5] = dz;
            }
            for (int a = 0; a < 3; ++a)
                for (int p = 0; p < 6; ++p) {
                    g[p] += w * J[a][p] * r[a];
                    for (int q = 0; q <= p; ++q) H[6 * p + q] += w * J[a][p] * J[a][q];
                }
        }
        if (it > 0 && cost > prev_cost && ++incr > prm.max_incr_cost) break;
        prev_cost = cost;
        for (int p = 0; p < 6; ++p) for (int q = p + 1; q < 6; ++q) H[6 * p + q] = H[6 * q + p];
        double d[6]; for (int p = 0; p < 6; ++p) d[p] = -g[p];
        if (!chol6(H, d)) { failed = true; return nit; }
        const Rot dR = rodrigues(d);
        const double nt[3] = { dR.m[0] * t[0] + dR.m[1] * t[1] + dR.m[2] * t[2] + d[3], dR.m[3] * t[0] + dR.m[4] * t[1] + dR.m[5] * t[2] + d[4],
                               dR.m[6] * t[0] + dR.m[7] * t[1] + dR.m[8] * t[2] + d[5] };
        R = mul(dR, R); t[0] = nt[0]; t[1] = nt[1]; t[2] = nt[2];
        ++nit;
        double m = 0; for (int p = 0; p < 6; ++p) m += d[p] * d[p];
        if (std::sqrt(m) < prm.min_mod_out_vector) break;
    }
    return nit;
}
}  // namespace

bool solveStereoChangeInPose(const double* pre_obs, const double* cur_obs, int n, const TStereoPoseCamera& cam, const double* initial,
                             const TStereoPoseSolverParams& prm, TStereoPoseResult& out)
{
    out = TStereoPoseResult();
    out.residual.assign(n > 0 ? n : 0, 1e30); out.inlier.assign(n > 0 ? n : 0, 0);
    if (n <= 0 || !pre_obs || !cur_obs) return false;
    Problem pb; pb.cur = cur_obs; pb.n = n; pb.cam = cam; pb.P.assign(3 * (size_t)n, 0.0); pb.ok.assign(n, 0);
    int nok = 0;
    for (int i = 0; i < n; ++i) {                                          // projectMatchTo3D, srba-stereo-slam_utils.h:558-574
        const double ul = pre_obs[3 * i], vl = pre_obs[3 * i + 1], ur = pre_obs[3 * i + 2];
        const double den = cam.fl * (cam.cur - ur) + cam.fr * (ul - cam.cul);
        if (!(den > 1e-9)) continue;                                       // zero / negative disparity: no 3-D point
        const double b_d = cam.baseline / den;
        pb.P[3 * i] = b_d * cam.fr * (ul - cam.cul); pb.P[3 * i + 1] = b_d * cam.fr * (vl - cam.cvl); pb.P[3 * i + 2] = b_d * cam.fl * cam.fr;
        pb.ok[i] = 1; ++nok;
    }
    const int min_pairs = prm.bad_tracking_th > 3 ? prm.bad_tracking_th : 3;
    if (nok < min_pairs) return false;
    double w0[3] = { 0, 0, 0 }, t[3] = { 0, 0, 0 };
    if (initial) { w0[0] = initial[0]; w0[1] = initial[1]; w0[2] = initial[2]; t[0] = initial[3]; t[1] = initial[4]; t[2] = initial[5]; }
    Rot R = rodrigues(w0);
    bool failed = false;
    out.num_it = run_stage(pb, pb.ok, prm.initial_max_iters, prm, R, t, failed);
    if (failed) return false;
    std::vector<uint8_t> active(n, 0); int nin = 0;
    for (int i = 0; i < n; ++i) {
        double r[3], Q[3];
        if (!pb.ok[i] || !residual_at(pb, i, R, t, r, Q)) continue;
        if (r[0] * r[0] + r[1] * r[1] + r[2] * r[2] <= prm.residual_threshold) { active[i] = 1; ++nin; }
    }
    if (nin < min_pairs) return false;
    out.num_it_final = run_stage(pb, active, prm.max_iters, prm, R, t, failed);
    if (failed) return false;
    for (int i = 0; i < n; ++i) {
        double r[3], Q[3];
        if (pb.ok[i] && residual_at(pb, i, R, t, r, Q)) out.residual[i] = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        out.inlier[i] = active[i];
    }
    rot_log(R, out.pose); out.pose[3] = t[0]; out.pose[4] = t[1]; out.pose[5] = t[2];
    out.num_inliers = nin;
    for (int p = 0; p < 6; ++p) if (!std::isfinite(out.pose[p])) return false;
    out.valid = true;
    return true;
}

extern "C" int srba_solve_change_in_pose(const double* pre_obs, const double* cur_obs, int n, const double* cam6, const double* initial6,
                                         const double* prm8, double* pose6, double* residual, uint8_t* inlier, int* its2)
{
    TStereoPoseCamera cam = { cam6[0], cam6[1], cam6[2], cam6[3], cam6[4], cam6[5] };
    TStereoPoseSolverParams p;
    p.initial_max_iters = (int)prm8[0]; p.max_iters = (int)prm8[1]; p.max_incr_cost = (int)prm8[2]; p.residual_threshold = prm8[3];
    p.min_mod_out_vector = prm8[4]; p.bad_tracking_th = (int)prm8[5]; p.use_robust_kernel = prm8[6] != 0; p.kernel_param = prm8[7];
    TStereoPoseResult r;
    const bool ok = solveStereoChangeInPose(pre_obs, cur_obs, n, cam, initial6, p, r);
    for (int i = 0; i < 6; ++i) pose6[i] = r.pose[i];
    for (int i = 0; i < n; ++i) { residual[i] = r.residual[i]; inlier[i] = r.inlier[i]; }
    its2[0] = r.num_it; its2[1] = r.num_it_final;
    return ok ? 1 : 0;
}
