// stereo_pose_solver.h -- robust least-squares change in pose between two stereo frames from tracked feature pairs.
// Stands in for the solver inside famoreno/stereo-vo behind rso::CStereoOdometryEstimator::processNewImagePair (outPose,
// reference CSRBAStereoSLAMEstimator.cpp:327-330) and ::getChangeInPose (:2139-2146, the data-association filter of
// :2113-2177).  stereo-vo is not vendored and carries no version pin, so the numerical recipe below is THIS repo's
// (parity unpinned); what the reference relies on is kept: the [LEAST_SQUARES] keys (kitti.ini:40-55), an initial guess
// [w1,w2,w3,tx,ty,tz] = pose of the PREVIOUS frame with respect to the CURRENT one (EST.cpp:2131-2135), two stages
// (num_it / num_it_final, EST.cpp:2151), one residual per tracked pair in input order (EST.cpp:2166-2176) and `valid`.
//
// Model: the previous frame's stereo observation (ul, vl, ur) is triangulated with the reference's own closed form
// (projectMatchTo3D, srba-stereo-slam_utils.h:558-574), moved by the pose (R, t) and projected into the current
// rectified pair; the residual is the 3-vector (ul, vl, ur) difference.  Gauss-Newton with pseudo-Huber IRLS weights on
// a left-multiplied se(3) increment; stage 1 runs on every pair, pairs whose squared residual exceeds
// `residual_threshold` are dropped, stage 2 refines on the rest.  out_residual = squared reprojection error in px^2.
#pragma once
#include <cstdint>
#include <vector>

struct TStereoPoseSolverParams {
    int initial_max_iters = 10, max_iters = 100, max_incr_cost = 3;       // kitti.ini:41-43 (30 / 30 / 3)
    double residual_threshold = 1.0, min_mod_out_vector = 0.001;          // kitti.ini:44-45 (15 / 0.001)
    int bad_tracking_th = 5;                                              // kitti.ini:47
    bool use_robust_kernel = true; double kernel_param = 3.0;             // kitti.ini:49-50 (true / 2)
};
struct TStereoPoseCamera { double cul, cvl, fl, cur, fr, baseline; };     // same six numbers projectMatchTo3D reads
struct TStereoPoseResult {
    double pose[6] = { 0, 0, 0, 0, 0, 0 };    // [w1,w2,w3,tx,ty,tz], previous frame wrt current frame
    bool valid = false; int num_it = 0, num_it_final = 0, num_inliers = 0;
    std::vector<double> residual;             // per pair, px^2 (1e30 for a pair that cannot be triangulated / projected)
    std::vector<uint8_t> inlier;              // per pair, after stage 1
};
// pre_obs / cur_obs: n x 3 doubles (ul, vl, ur).  initial may be NULL (identity).
bool solveStereoChangeInPose(const double* pre_obs, const double* cur_obs, int n, const TStereoPoseCamera& cam, const double* initial,
                             const TStereoPoseSolverParams& prm, TStereoPoseResult& out);
