"""Oracle restatement (plain Python / numpy) of the geometric data-association filters -- TEST INFRASTRUCTURE ONLY.

  find_fundamental_ransac   cv::findFundamentalMat(p1, p2, FM_RANSAC, param1, param2, mask) as OpenCV 2.4 runs it (the call
                            at CSRBAStereoSLAMEstimator.cpp:2043): RANSAC over 7-point sets, cv::RNG seeded with -1,
                            <= 2000 iterations, no final refit, >= 15 points.  OpenCV is not vendored by the reference; the
                            algorithm is the published one (calib3d fundam.cpp / modelest.cpp).  PINNED against cv2 4.13's
                            findFundamentalMat inlier masks (tests/golden/fundmat.npz, oracle/tools/make_golden.py).
  solve_change_in_pose      the two-stage robust Gauss-Newton behind rso::CStereoOdometryEstimator::getChangeInPose
                            (CSRBAStereoSLAMEstimator.cpp:2139-2146).  stereo-vo is not vendored and has no version pin:
                            PARITY UNPINNED; this is the recipe srba-stereo-slam_b200/host/stereo_pose_solver.h documents.
Scalar arithmetic is written out in the same order as the C++ so both sides round identically.
"""
import math
import numpy as np

DBL_EPSILON = 2.220446049250313e-16
DBL_MIN = 2.2250738585072014e-308
FLT_EPSILON = 1.1920928955078125e-07


class CvRng:
    """cv::RNG: multiply-with-carry, state = (uint32)state * 4164903690 + (state >> 32)."""

    def __init__(self, seed=0xFFFFFFFFFFFFFFFF):
        self.state = seed if seed else 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF


def _nullspace_7x9(A):
    """Last two columns of Q in the Householder QR of A^T (9x7); reflectors map onto -sign(x0)|x| e0."""
    M = [[A[i][j] for i in range(7)] for j in range(9)]
    V = [[0.0] * 9 for _ in range(7)]
    beta = [0.0] * 7
    for k in range(7):
        nrm = 0.0
        for i in range(k, 9):
            nrm += M[i][k] * M[i][k]
        nrm = math.sqrt(nrm)
        if nrm == 0.0:
            continue
        alpha = -nrm if M[k][k] >= 0 else nrm
        for i in range(k, 9):
            V[k][i] = M[i][k]
        V[k][k] -= alpha
        vn = 0.0
        for i in range(k, 9):
            vn += V[k][i] * V[k][i]
        beta[k] = 2.0 / vn if vn > 0 else 0.0
        for j in range(k, 7):
            s = 0.0
            for i in range(k, 9):
                s += V[k][i] * M[i][j]
            s *= beta[k]
            for i in range(k, 9):
                M[i][j] -= s * V[k][i]
    out = []
    for c in range(2):
        q = [1.0 if i == 7 + c else 0.0 for i in range(9)]
        for k in range(6, -1, -1):
            s = 0.0
            for i in range(k, 9):
                s += V[k][i] * q[i]
            s *= beta[k]
            for i in range(k, 9):
                q[i] -= s * V[k][i]
        out.append(q)
    return out[0], out[1]


def _det3(m):
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6])


def _solve_cubic(c):
    a0, b0, c0, d0 = c
    if a0 == 0:
        if b0 == 0:
            if c0 == 0:
                return []
            return [-d0 / c0]
        d = c0 * c0 - 4 * b0 * d0
        if d < 0:
            return []
        d = math.sqrt(d)
        q1 = (-c0 + d) * 0.5; q2 = (-c0 - d) * 0.5
        r = [q1 / b0, d0 / q1] if abs(q1) > abs(q2) else [q2 / b0, d0 / q2]
        return r if d > 0 else r[:1]
    a1 = b0 / a0; a2 = c0 / a0; a3 = d0 / a0
    Q = (a1 * a1 - 3 * a2) * (1. / 9); R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54)
    Qcubed = Q * Q * Q
    d = Qcubed - R * R
    if d >= 0:
        if Qcubed == 0:
            return [-a1 / 3] * 3
        theta = math.acos(max(-1.0, min(1.0, R / math.sqrt(Qcubed))))
        t0 = -2 * math.sqrt(Q); t1 = theta * (1. / 3); t2 = a1 * (1. / 3)
        return [t0 * math.cos(t1) - t2, t0 * math.cos(t1 + (2. * math.pi / 3)) - t2, t0 * math.cos(t1 + (4. * math.pi / 3)) - t2]
    d = math.sqrt(-d)
    e = math.pow(d + abs(R), 0.333333333333)
    if R > 0:
        e = -e
    return [(e + Q / e) - a1 * (1. / 3)]


def fundamental_7point(m1, m2):
    """Up to three 3x3 matrices F (row-major lists of 9) with (m2,1)^T F (m1,1) = 0 for the 7 correspondences."""
    A = []
    for i in range(7):
        x0, y0 = m1[i]; x1, y1 = m2[i]
        A.append([x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, 1.0])
    g1, g2 = _nullspace_7x9(A)
    dlt = [g1[i] - g2[i] for i in range(9)]

    def det_at(l):
        return _det3([g2[i] + l * dlt[i] for i in range(9)])
    p0, p1, pm, p2 = det_at(0.0), det_at(1.0), det_at(-1.0), det_at(2.0)
    c3 = p0
    c1 = 0.5 * (p1 + pm) - p0
    s = 0.5 * (p1 - pm)
    c0 = (p2 - 4 * c1 - c3 - 2 * s) / 6.0
    c2 = s - c0
    out = []
    for r in _solve_cubic([c0, c1, c2, c3]):
        lam, mu = r, 1.0
        w = dlt[8] * r + g2[8]
        F = [0.0] * 9
        if abs(w) > DBL_EPSILON:
            mu = 1.0 / w; lam *= mu; F[8] = 1.0
        for i in range(8):
            F[i] = dlt[i] * lam + g2[i] * mu
        out.append(F)
    return out


def _newest_point_ok(pts, count):
    if count <= 2:
        return True
    i = count - 1
    for j in range(i):
        dx1 = pts[j][0] - pts[i][0]; dy1 = pts[j][1] - pts[i][1]
        for k in range(j):
            dx2 = pts[k][0] - pts[i][0]; dy2 = pts[k][1] - pts[i][1]
            if abs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return False
    return True


def _draw_subset(rng, m1, m2, count, max_attempts=300):
    s1 = [None] * 7; s2 = [None] * 7; idx = [0] * 7
    iters = 0; i = 0
    while iters < max_attempts:
        i = 0
        while i < 7 and iters < max_attempts:
            idx[i] = rng.next() % count
            if idx[i] in idx[:i]:
                continue
            s1[i] = m1[idx[i]]; s2[i] = m2[idx[i]]
            if not _newest_point_ok(s1, i + 1) or not _newest_point_ok(s2, i + 1):
                iters += 1
                continue
            i += 1
        break
    return (i == 7 and iters < max_attempts), s1, s2


def _draw_subset_cv4(rng, m1, m2, count, max_attempts=10000):
    """OpenCV >= 3: draw all seven (duplicates redrawn), then one collinearity check of the last point; redraw everything on failure."""
    for _ in range(max_attempts):
        idx = []
        for i in range(7):
            while True:
                v = rng.next() % count
                if v not in idx:
                    break
            idx.append(v)
        s1 = [m1[i] for i in idx]; s2 = [m2[i] for i in idx]
        if _newest_point_ok(s1, 7) and _newest_point_ok(s2, 7):
            return True, s1, s2
    return False, None, None


def _errors(m1, m2, F):
    x1, y1, x2, y2 = m1[:, 0], m1[:, 1], m2[:, 0], m2[:, 1]
    a = F[0] * x1 + F[1] * y1 + F[2]; b = F[3] * x1 + F[4] * y1 + F[5]; c = F[6] * x1 + F[7] * y1 + F[8]
    s2 = 1. / (a * a + b * b); d2 = x2 * a + y2 * b + c
    a = F[0] * x2 + F[3] * y2 + F[6]; b = F[1] * x2 + F[4] * y2 + F[7]; c = F[2] * x2 + F[5] * y2 + F[8]
    s1 = 1. / (a * a + b * b); d1 = x1 * a + y1 * b + c
    return np.maximum(d1 * d1 * s1, d2 * d2 * s2).astype(np.float32)


def _update_iters(p, ep, model_points, max_iters):
    p = min(max(p, 0.), 1.); ep = min(max(ep, 0.), 1.)
    num = max(1. - p, DBL_MIN); denom = 1. - math.pow(1. - ep, model_points)
    if denom < DBL_MIN:
        return 0
    num = math.log(num); denom = math.log(denom)
    return max_iters if (denom >= 0 or -num >= max_iters * (-denom)) else int(np.rint(num / denom))


def find_fundamental_ransac(p1, p2, param1=3.0, param2=0.99, cv4=False):
    """p1, p2: [n, 2] float32.  Returns (mask uint8 [n], F [9] or None, iterations).  cv4=False: OpenCV 2.4 (the reference's
    dependency); cv4=True: the OpenCV >= 3 variant (1000 iterations, whole-sample subset check) that cv2 4.13 pins exactly."""
    n = len(p1)
    mask = np.zeros(n, np.uint8)
    if n < 15:
        return mask, None, 0
    if param1 <= 0:
        param1 = 3
    if param2 < DBL_EPSILON or param2 > 1 - DBL_EPSILON:
        param2 = 0.99
    m1 = np.asarray(p1, np.float32).astype(np.float64); m2 = np.asarray(p2, np.float32).astype(np.float64)
    l1 = [(float(a), float(b)) for a, b in m1]; l2 = [(float(a), float(b)) for a, b in m2]
    thr2 = param1 * param1
    rng = CvRng()
    niters, max_good, best_f, it = (1000 if cv4 else 2000), 0, None, 0
    while it < niters:
        ok, s1, s2 = _draw_subset_cv4(rng, l1, l2, n) if cv4 else _draw_subset(rng, l1, l2, n)
        if not ok:
            if it == 0:
                return mask, None, 0
            break
        for F in fundamental_7point(s1, s2):
            tm = _errors(m1, m2, F) <= thr2
            good = int(tm.sum())
            if good > max(max_good, 6):
                mask = tm.astype(np.uint8); best_f = F; max_good = good
                niters = _update_iters(param2, (n - good) / n, 7, niters)
        it += 1
    if max_good <= 0:
        return np.zeros(n, np.uint8), None, it
    return mask, best_f, it


# ---- change in pose ----------------------------------------------------------------------------------------------------
def _rodrigues(w):
    th = math.sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2])
    if th < 1e-8:
        a = 1.0 - th * th / 6.0; b = 0.5 - th * th / 24.0
    else:
        a = math.sin(th) / th; b = (1.0 - math.cos(th)) / (th * th)
    x, y, z = w
    return np.array([[1 - b * (y * y + z * z), -a * z + b * x * y, a * y + b * x * z],
                     [a * z + b * x * y, 1 - b * (x * x + z * z), -a * x + b * y * z],
                     [-a * y + b * x * z, a * x + b * y * z, 1 - b * (x * x + y * y)]])


def _rot_log(R):
    c = 0.5 * (R[0, 0] + R[1, 1] + R[2, 2] - 1.0)
    c = max(-1.0, min(1.0, c))
    th = math.acos(c)
    k = 0.5 if th < 1e-8 else th / (2.0 * math.sin(th))
    return [k * (R[2, 1] - R[1, 2]), k * (R[0, 2] - R[2, 0]), k * (R[1, 0] - R[0, 1])]


def _residuals(P, cur, cam, R, t):
    cul, cvl, fl, cur_c, fr, base = cam
    Q = P @ R.T + t
    ok = Q[:, 2] > 1e-6
    iz = 1.0 / np.where(ok, Q[:, 2], 1.0)
    r = np.stack([fl * Q[:, 0] * iz + cul - cur[:, 0], fl * Q[:, 1] * iz + cvl - cur[:, 1], fr * (Q[:, 0] - base) * iz + cur_c - cur[:, 2]], 1)
    return r, Q, ok


def _stage(P, cur, cam, active, max_iters, prm, R, t):
    cul, cvl, fl, cur_c, fr, base = cam
    k2 = prm["kernel_param"] ** 2
    incr, nit, prev_cost = 0, 0, 0.0
    for it in range(max_iters):
        r, Q, ok = _residuals(P, cur, cam, R, t)
        use = active & ok
        rr, QQ = r[use], Q[use]
        e2 = (rr * rr).sum(1)
        if prm["use_robust_kernel"]:
            s = np.sqrt(1.0 + e2 / k2); w = 1.0 / s; cost = float((2.0 * k2 * (s - 1.0)).sum())
        else:
            w = np.ones(len(e2)); cost = float(e2.sum())
        iz = 1.0 / QQ[:, 2]; xb = QQ[:, 0] - base
        z = np.zeros(len(iz))
        dq = np.stack([np.stack([fl * iz, z, -fl * QQ[:, 0] * iz * iz], 1), np.stack([z, fl * iz, -fl * QQ[:, 1] * iz * iz], 1),
                       np.stack([fr * iz, z, -fr * xb * iz * iz], 1)], 1)                 # [n, 3 (residual), 3 (dQ)]
        dx, dy, dz = dq[:, :, 0], dq[:, :, 1], dq[:, :, 2]
        Qx, Qy, Qz = QQ[:, 0:1], QQ[:, 1:2], QQ[:, 2:3]
        J = np.stack([dz * Qy - dy * Qz, dx * Qz - dz * Qx, dy * Qx - dx * Qy, dx, dy, dz], 2)   # [n, 3, 6]
        if it > 0 and cost > prev_cost:
            incr += 1
            if incr > prm["max_incr_cost"]:
                break
        prev_cost = cost
        H = np.einsum("n,nap,naq->pq", w, J, J); g = np.einsum("n,nap,na->p", w, J, rr)
        try:
            Lc = np.linalg.cholesky(H)
        except np.linalg.LinAlgError:
            return nit, R, t, True
        if np.min(np.diag(Lc)) ** 2 <= 1e-12:
            return nit, R, t, True
        d = np.linalg.solve(H, -g)
        dR = _rodrigues(d[:3])
        t = dR @ t + d[3:]
        R = dR @ R
        nit += 1
        if math.sqrt(float(d @ d)) < prm["min_mod_out_vector"]:
            break
    return nit, R, t, False


DEFAULT_LS = dict(initial_max_iters=10, max_iters=100, max_incr_cost=3, residual_threshold=1.0, min_mod_out_vector=0.001,
                  bad_tracking_th=5, use_robust_kernel=True, kernel_param=3.0)


def solve_change_in_pose(pre_obs, cur_obs, cam, initial=None, prm=None):
    """pre_obs, cur_obs: [n, 3] (ul, vl, ur); cam = (cul, cvl, fl, cur, fr, baseline).  Returns dict(valid, pose[6] =
    [w, t] of the previous frame wrt the current one, residual[n] in px^2 (1e30 = unusable pair), inlier[n], num_it, num_it_final)."""
    p = dict(DEFAULT_LS); p.update(prm or {})
    pre = np.asarray(pre_obs, np.float64).reshape(-1, 3); cur = np.asarray(cur_obs, np.float64).reshape(-1, 3)
    n = len(pre)
    out = dict(valid=False, pose=[0.0] * 6, residual=np.full(n, 1e30), inlier=np.zeros(n, np.uint8), num_it=0, num_it_final=0)
    if n == 0:
        return out
    cul, cvl, fl, cur_c, fr, base = cam
    den = fl * (cur_c - pre[:, 2]) + fr * (pre[:, 0] - cul)
    okp = den > 1e-9
    b_d = base / np.where(okp, den, 1.0)
    P = np.stack([b_d * fr * (pre[:, 0] - cul), b_d * fr * (pre[:, 1] - cvl), b_d * fl * fr], 1)
    min_pairs = max(p["bad_tracking_th"], 3)
    if int(okp.sum()) < min_pairs:
        return out
    w0 = list(initial[:3]) if initial is not None else [0.0, 0.0, 0.0]
    t = np.array(initial[3:6], np.float64) if initial is not None else np.zeros(3)
    R = _rodrigues(w0)
    out["num_it"], R, t, failed = _stage(P, cur, cam, okp, p["initial_max_iters"], p, R, t)
    if failed:
        return out
    r, Q, ok = _residuals(P, cur, cam, R, t)
    active = okp & ok & ((r * r).sum(1) <= p["residual_threshold"])
    if int(active.sum()) < min_pairs:
        return out
    out["num_it_final"], R, t, failed = _stage(P, cur, cam, active, p["max_iters"], p, R, t)
    if failed:
        return out
    r, Q, ok = _residuals(P, cur, cam, R, t)
    good = okp & ok
    out["residual"] = np.where(good, (r * r).sum(1), 1e30)
    out["inlier"] = active.astype(np.uint8)
    out["pose"] = _rot_log(R) + [float(t[0]), float(t[1]), float(t[2])]
    out["valid"] = bool(np.all(np.isfinite(out["pose"])))
    return out


# ---- SE(3) helpers shared with the sequence oracle (mrpt::poses::CPose3DRotVec / CPose3D semantics) -------------------
def rv_compose(a, b):
    """(w, t) poses as 6-lists [wx, wy, wz, x, y, z]: a (+) b."""
    Ra, Rb = _rodrigues(a[:3]), _rodrigues(b[:3])
    t = np.asarray(a[3:6]) + Ra @ np.asarray(b[3:6])
    return rot_log_full(Ra @ Rb) + [float(t[0]), float(t[1]), float(t[2])]


def rv_inverse(a):
    R = _rodrigues(a[:3])
    t = -(R.T @ np.asarray(a[3:6]))
    return [-a[0], -a[1], -a[2], float(t[0]), float(t[1]), float(t[2])]


def rot_log_full(R):
    """mrpt_lite.h rot_log (handles angles near pi, which CPose3D(-90, 0, -90 deg) style sensor poses reach)."""
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    c = max(-1.0, min(1.0, 0.5 * (tr - 1.0)))
    th = math.acos(c)
    ax, ay, az = R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]
    if th < 1e-8:
        return [0.5 * ax, 0.5 * ay, 0.5 * az]
    if math.pi - th < 1e-6:
        d = [R[0, 0], R[1, 1], R[2, 2]]
        i = (0 if d[0] > d[2] else 2) if d[0] > d[1] else (1 if d[1] > d[2] else 2)
        u = [0.0, 0.0, 0.0]
        u[i] = math.sqrt(max(0.0, 0.5 * (d[i] + 1.0)))
        for j in range(3):
            if j != i:
                u[j] = (R[i, j] + R[j, i]) / (4.0 * u[i])
        sgn = -1.0 if (ax * u[0] + ay * u[1] + az * u[2]) < 0 else 1.0
        return [sgn * th * u[j] for j in range(3)]
    k = th / (2.0 * math.sin(th))
    return [k * ax, k * ay, k * az]


def ypr_to_rot(yaw, pitch, roll):
    cy, sy, cp, sp, cr, sr = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch), math.cos(roll), math.sin(roll)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])
