"""Oracle restatement (plain Python / numpy) of the host-side data-association steps downstream of the keyframe
match -- TEST INFRASTRUCTURE ONLY.  Each function follows the reference line by line:
  filter_direction     CSRBAStereoSLAMEstimator.cpp:1883-1946
  filter_orb_distance  CSRBAStereoSLAMEstimator.cpp:1951-2010 (with the `= ++stage1_counter` bug at :2003-2007)
  assign_match_ids     CSRBAStereoSLAMEstimator.cpp:572-617
  project_match_to_3d  srba-stereo-slam_utils.h:558-574
The reference has no tests or golden vectors for these; parity is by construction (the source is in /root/reference).
"""
import math
import numpy as np

TRACKED, NON_TRACKED, REJ_SLOPE, REJ_ORB, REJ_FUND, REJ_POSE, REJ_CONSISTENCY = range(7)


def filter_direction(this_xy, other_xy, offset, status):
    status = list(status)
    n_bins, bin_width = 36, 10
    hist = [0] * n_bins
    hist_m = [[] for _ in range(n_bins)]
    for k in range(len(status)):
        if status[k] != TRACKED:
            continue
        tx, ty = np.float32(this_xy[k][0]), np.float32(this_xy[k][1])
        ox, oy = np.float32(other_xy[k][0]), np.float32(np.float32(other_xy[k][1]) + np.float32(offset))
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.float32(oy - ty) / np.float32(ox - tx)
        ang = 180.0 / math.pi * float(np.arctan(np.float32(r), dtype=np.float32)) + 180.0
        if not (0.0 <= ang < 360.0):
            continue
        b = int(math.floor(ang / bin_width))
        hist[b] += 1; hist_m[b].append(k)
    mode, best = 0, 0
    for k in range(n_bins):
        if hist[k] > best:
            best, mode = hist[k], k
    keep = {(mode - 1) % n_bins, mode, (mode + 1) % n_bins}
    for k in range(n_bins):
        if k in keep:
            continue
        for idx in hist_m[k]:
            status[idx] = REJ_SLOPE
    return status


def filter_orb_distance(query_idx, train_idx, distance, n_other, max_dist, status):
    status = list(status)
    other_first = [-1] * n_other
    other_second = [float(np.finfo(np.float32).max)] * n_other
    counter = 0
    for q, t, d in zip(query_idx, train_idx, distance):
        d = float(np.float32(d))
        if status[q] != TRACKED:
            continue
        if d > max_dist:
            status[q] = REJ_ORB; continue
        if d > other_second[t]:
            status[q] = NON_TRACKED; continue
        if other_first[t] != -1:
            status[other_first[t]] = NON_TRACKED; counter -= 1
        other_first[t] = q
        counter += 1
        other_second[t] = float(np.float32(counter))          # the bug: the running counter replaces the distance
    return status


def assign_match_ids(n_matches, da_infos, sorted_idx, kf_ids, last_match_id):
    """da_infos: list of (kf_idx, tracking_first[n_matches] with -1 = invalid); kf_ids: list of id lists per keyframe."""
    out = [0] * n_matches
    found = set(); tracked_n = new_n = 0
    for m in range(n_matches):
        tracked = False
        for k in sorted_idx:
            kf_idx, first = da_infos[k]
            if first[m] != -1:
                oid = kf_ids[kf_idx][first[m]]
                if oid in found:
                    break
                found.add(oid); out[m] = oid; tracked_n += 1; tracked = True
                break
        if not tracked:
            last_match_id += 1; out[m] = last_match_id; new_n += 1
    return out, last_match_id, tracked_n, new_n


def project_match_to_3d(ul, vl, ur, cul, cvl, fl, cur, fr, baseline):
    b_d = baseline / (fl * (cur - ur) + fr * (ul - cul))
    return b_d * fr * (ul - cul), b_d * fr * (vl - cvl), b_d * fl * fr
