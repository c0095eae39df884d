"""Oracle chain for a whole sequence -- TEST INFRASTRUCTURE ONLY.

A Python restatement of what the reference does between a stereo frame arriving and an observation list reaching
rba.define_new_keyframe (CSRBAStereoSLAMEstimator.cpp:29-926), stacked on the per-stage oracles:
  extraction + L/R match + BF match + BoW     oracle/oracle.py   (frontend_oracle.c; cv2-pinned where cv2 can pin)
  direction / ORB-distance filters, id rules  oracle/da_oracle.py
  F-matrix RANSAC, change-in-pose solver      oracle/geometry_oracle.py
  inter-frame id propagation (ifmDescBF)      StereoOdometryOracle below (stereo-vo absent: this repo's convention)
  keyframe decisions                          SlamOracle below (EST.cpp:263-315 adaptive thresholds, :359-387 triggers,
                                              :413-562 check, :572-617 ids, :666-699 observations, :1737-1878 similar KFs)
  SRBA stand-in                               SubmapGraph below = host/stereo_slam_frontend.h CSubmapGraphBackend (SRBA absent)
tests/test_gpu_sequence.py runs the C++ adapters + CUDA library on the same frames and diffs the observation lists.
"""
import math
from collections import deque
import numpy as np

from . import oracle as orc
from . import da_oracle as dao
from . import geometry_oracle as geo

INF = (1 << 64) - 1
INVALID = -1


class StereoOdometryOracle:
    """rso::CStereoOdometryEstimator as host/stereo_odometry_b200.cpp implements it."""

    def __init__(self, cam, nfeatures, nlevels, fast_th=20, orb_th=60, fast_min=5, fast_max=30, orb_min=30, orb_max=100,
                 max_y_diff=2.0, ls=None, use_previous_pose_as_initial=True):
        self.cam = cam; self.nf = nfeatures; self.nl = nlevels
        self.fast_th = fast_th; self.orb_th = orb_th; self.orb_default = orb_th
        self.fast_min, self.fast_max, self.orb_min, self.orb_max = fast_min, fast_max, orb_min, orb_max
        self.max_y_diff = max_y_diff; self.ls = ls or {}; self.use_prev = use_previous_pose_as_initial
        self.have = False; self.next_id = 0; self.ids = []; self.kf_ids = []
        self.prev = None; self.cur = None; self.last_guess = [0.0] * 6; self.guess0 = [0.0] * 6; self.next_id0 = 0

    def set_fast(self, th): self.fast_th = max(self.fast_min, min(th, self.fast_max))
    def set_orb(self, th): self.orb_th = max(self.orb_min, min(th, self.orb_max))
    def fast_is_min(self): return self.fast_th <= self.fast_min
    def orb_is_max(self): return self.orb_th >= self.orb_max

    def process(self, left, right, repeat=False):
        if not repeat:
            self.images = (left, right)
            if self.have:
                self.prev = self.cur
            self.next_id0 = self.next_id; self.guess0 = list(self.last_guess)
        else:
            left, right = self.images
            self.next_id = self.next_id0
        kl, dl, _ = orc.orb_extract(left, nfeatures=self.nf, nlevels=self.nl, fast_threshold=self.fast_th)
        kr, dr, _ = orc.orb_extract(right, nfeatures=self.nf, nlevels=self.nl, fast_threshold=self.fast_th)
        m = orc.match_rowband(kl, dl, kr, dr, max_y_diff=self.max_y_diff, max_dist=self.orb_th)
        nm = len(m)
        desc = dl[m["queryIdx"]] if nm else np.zeros((0, 32), np.uint8)
        ids = [0] * nm; tracked = 0; tracked_kf = 0; pairs = []
        kfset = set(self.kf_ids)
        if self.prev is not None and nm and len(self.prev["desc"]):
            bf = orc.match_bf(desc, self.prev["desc"]); claimed = set()
            for k in range(nm):
                t = int(bf["trainIdx"][k])
                if bf["distance"][k] <= self.orb_default_if() and t not in claimed:
                    ids[k] = self.prev["ids"][t]; claimed.add(t); tracked += 1; tracked_kf += ids[k] in kfset
                    pairs.append((t, k))
                else:
                    ids[k] = self.next_id; self.next_id += 1
        else:
            for k in range(nm):
                ids[k] = self.next_id; self.next_id += 1
        self.ids = ids
        self.cur = dict(kl=kl, kr=kr, dl=dl, dr=dr, m=m, desc=desc, ids=ids)
        res = dict(n_left=len(kl), n_right=len(kr), stereo_matches=nm, tracked=tracked, tracked_kf=tracked_kf, valid=nm >= 8,
                   out_pose=[0.0] * 6)
        if self.have and res["valid"]:
            r = self.change_in_pose(pairs, self.prev, self.cur, self.guess0 if self.use_prev else None)
            res["valid"] = r["valid"]; res["out_pose"] = geo.rv_inverse(r["pose"])
            if r["valid"]:
                self.last_guess = list(r["pose"])
        self.have = True
        return res

    def orb_default_if(self):
        return self.orb_default            # params_if_match.orb_max_distance: the configured value, not the adapted threshold

    def change_in_pose(self, pairs, pre, cur, initial):
        def obs(kf, idx):
            q = kf["m"]["queryIdx"][idx]; t = kf["m"]["trainIdx"][idx]
            return np.stack([kf["kl"]["x"][q], kf["kl"]["y"][q], kf["kr"]["x"][t]], 1).astype(np.float64)
        if not pairs:
            return dict(valid=False, pose=[0.0] * 6, residual=np.zeros(0), inlier=np.zeros(0, np.uint8), num_it=0, num_it_final=0)
        a = np.array([p[0] for p in pairs]); b = np.array([p[1] for p in pairs])
        return geo.solve_change_in_pose(obs(pre, a), obs(cur, b), self.cam, initial, self.ls)

    def get_values(self):
        c = self.cur
        return dict(kl=c["kl"], kr=c["kr"], dl=c["dl"], dr=c["dr"], m=c["m"], ids=list(c["ids"]))

    def reset_ids(self): self.kf_ids = list(self.ids)
    set_this_frame_as_kf = reset_ids


class SubmapGraph:
    """CSubmapGraphBackend (host/stereo_slam_frontend.cpp)."""

    def __init__(self, submap, tree_depth, opt_depth, sensor_rv):
        self.submap = max(submap, 1); self.tree = tree_depth; self.opt = opt_depth
        self.K = sensor_rv; self.Kinv = geo.rv_inverse(sensor_rv)
        self.pose = []; self.adj = []; self.base_kf = {}

    def bfs(self, a, b, max_depth):
        if a >= len(self.adj) or b >= len(self.adj):
            return INF
        if a == b:
            return 0
        dist = {a: 0}; q = deque([a])
        while q:
            u = q.popleft()
            if dist[u] >= max_depth:
                continue
            for v in self.adj[u]:
                if v not in dist:
                    dist[v] = dist[u] + 1
                    if v == b:
                        return dist[v]
                    q.append(v)
        return INF

    def define_new_keyframe(self, obs_ids, initial_rv):
        kid = len(self.pose); self.adj.append([])
        centre = lambda k: self.submap * (k // self.submap)
        created = []

        def link(a, b):
            self.adj[a].append(b); self.adj[b].append(a); created.append((a, b))
        if kid == 0:
            self.pose.append([0.0] * 6)
        else:
            step = geo.rv_compose(geo.rv_compose(self.K, initial_rv), self.Kinv)
            self.pose.append(geo.rv_compose(self.pose[kid - 1], step))
            if kid % self.submap != 0:
                link(centre(kid), kid)
            else:
                votes = {}
                for i in obs_ids:
                    if i in self.base_kf:
                        c = centre(self.base_kf[i]); votes[c] = votes.get(c, 0) + 1
                for c, n in sorted(sorted(votes.items()), key=lambda kv: -kv[1]):
                    if self.bfs(kid, c, self.tree) >= self.opt and n >= 4:
                        link(c, kid)
                if not created:
                    link(centre(kid - 1), kid)
        for i in obs_ids:
            self.base_kf.setdefault(i, kid)
        return kid

    def topo(self, a, b): return self.bfs(a, b, self.tree)

    def tree_pose(self, root, kid):
        if self.bfs(root, kid, self.tree) == INF:
            return None
        return geo.rv_compose(geo.rv_inverse(self.pose[root]), self.pose[kid])


def _norm3(v): return math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


class SlamOracle:
    def __init__(self, voc, opt, cam, image_rows, ls=None, camera_pose_on_robot=(0, 0, 0, 0, 0, 0)):
        """opt: dict with the TSRBAStereoSLAMOptions fields (reference names); cam = (cul, cvl, fl, cur, fr, baseline)."""
        self.voc = voc; self.o = opt; self.cam = cam; self.rows = image_rows
        p = camera_pose_on_robot
        d2r = lambda x: x * math.pi / 180.0                      # DEG2RAD as the C++ macro rounds it
        Rc = geo.ypr_to_rot(d2r(p[3]), d2r(p[4]), d2r(p[5]))
        Ri = geo.ypr_to_rot(d2r(-90), 0.0, d2r(-90))
        self.K = geo.rot_log_full(Rc @ Ri) + [float(p[0]), float(p[1]), float(p[2])]
        self.Kinv = geo.rv_inverse(self.K)
        self.vo = StereoOdometryOracle(cam, opt["n_feats"], opt.get("n_levels", 1), fast_th=opt["detect_fast_th"], orb_th=opt.get("orb_max_distance", 60),
                                       max_y_diff=opt.get("max_y_diff", 2.0), ls=ls, use_previous_pose_as_initial=opt.get("use_previous_pose_as_initial", True))
        self.rba = SubmapGraph(opt["srba_submap_size"], opt["srba_max_tree_depth"], opt["srba_max_optimize_depth"], self.K)
        self.kfs = []; self.db = []
        self.last_kf_id = 0; self.last_match_id = 0
        self.max_rot_limit = opt["max_rotation"]; self.max_tr_limit = opt["max_translation"]
        self.max_rot = 2 * self.max_rot_limit; self.max_tr = 2 * self.max_tr_limit
        self.cur_pose = list(self.K); self.last_kf_pose = [0.0] * 6
        self.incr_kf = [0.0] * 6; self.incr_check = [0.0] * 6
        self.obs_lists = []; self.kf_frames = []; self.checks = []; self.frame_poses = []
        self.count = 0

    # ---- pieces ---------------------------------------------------------------------------------------------------------
    def _observations(self, kf):
        out = []
        RK = geo._rodrigues(self.K[:3]); tK = np.asarray(self.K[3:6])
        for i in range(len(kf["m"])):
            q, t = int(kf["m"]["queryIdx"][i]), int(kf["m"]["trainIdx"][i])
            lx, ly = kf["kl"]["x"][q], kf["kl"]["y"][q]; rx, ry = kf["kr"]["x"][t], kf["kr"]["y"][t]
            X, Y, Z = dao.project_match_to_3d(float(lx), float(ly), float(rx), *self.cam)
            P = [float(tK[a]) + float(RK[a, 0]) * X + float(RK[a, 1]) * Y + float(RK[a, 2]) * Z for a in range(3)]      # same order as CPose3DRotVec + TPoint3D
            out.append((int(kf["ids"][i]), float(lx), float(ly), float(rx), float(ry), P[0], P[1], P[2]))
        return out

    def _bow_insert(self, kf):
        w, wt = orc.bow_transform(self.voc, kf["dl"]); self.db.append(orc.bow_vector(self.voc, w, wt))

    def _similar_kfs(self, new_id, q):
        o = self.o
        with_pose = o["da_stage2_method"] in (2, 3)
        to_prev = geo.rv_inverse(self.incr_kf)
        sim = [new_id - 1]; poses = []
        if len(q) == 1:
            return sim, ([self._via_cpose3d(to_prev)] if with_pose else []), INVALID, False
        lc_id, potential = INVALID, False
        base = o["srba_submap_size"] * ((new_id - 1) // o["srba_submap_size"])
        th = 0.8 * q[0][1]
        for eid, score in q:
            if eid == new_id - 1:
                continue
            if score > th:
                sim.append(eid); continue
            topo = 0 if base == eid else self.rba.topo(base, eid)
            if lc_id == INVALID and topo > o["lc_distance"] and score > 0.05:
                lc_id = eid; potential = True; sim.append(eid)
        poses = [[0.0] * 6 for _ in sim]
        if with_pose:
            for k, s in enumerate(sim):
                if s == new_id - 1:
                    poses[k] = self._via_cpose3d(to_prev); continue
                tp = self.rba.tree_pose(new_id - 1, s)
                if tp is not None:
                    aux = geo.rv_compose(geo.rv_compose(self.Kinv, self._via_cpose3d(tp)), self.K)
                    poses[k] = self._via_cpose3d(geo.rv_compose(to_prev, aux))
        return sim, poses, lc_id, potential

    @staticmethod
    def _via_cpose3d(rv):
        """CPose3D <-> CPose3DRotVec round trip (rotation matrix in between): exp then log."""
        return geo.rot_log_full(geo._rodrigues(rv[:3])) + list(rv[3:6])

    def _internal_da(self, this_kf, other_kf, ini_pose):
        o = self.o
        n = len(this_kf["m"])
        info = dict(kf_idx=other_kf["id"], tracked=0, first=[-1] * n, dist=[0.0] * n)
        tq = this_kf["dl"][this_kf["m"]["queryIdx"]]; ot = other_kf["dl"][other_kf["m"]["queryIdx"]]
        bf = orc.match_bf(tq, ot)
        if len(bf) != n:
            return info
        status = [dao.TRACKED] * n
        txy = np.stack([this_kf["kl"]["x"][this_kf["m"]["queryIdx"]], this_kf["kl"]["y"][this_kf["m"]["queryIdx"]]], 1)
        oq = other_kf["m"]["queryIdx"][bf["trainIdx"]]
        oxy = np.stack([other_kf["kl"]["x"][oq], other_kf["kl"]["y"][oq]], 1)
        invalid = False
        if o["da_filter_by_direction"]:
            status = dao.filter_direction(txy, oxy, self.rows, status)
        if o["da_filter_by_orb_distance"]:
            status = dao.filter_orb_distance(bf["queryIdx"], bf["trainIdx"], bf["distance"], len(other_kf["m"]), o["max_orb_distance_da"], status)
        inl = sum(s == dao.TRACKED for s in status)
        if o["da_filter_by_fund_matrix"]:
            if inl < 15:
                invalid = True
            else:
                sel = [k for k in range(n) if status[k] == dao.TRACKED]
                mask, _, _ = geo.find_fundamental_ransac(oxy[sel].astype(np.float32), txy[sel].astype(np.float32), o["max_y_diff_epipolar"], o["ransac_fit_prob"])
                for j, k in enumerate(sel):
                    if mask[j] == 0:
                        status[k] = dao.REJ_FUND
                inl = sum(s == dao.TRACKED for s in status)
        if o["da_filter_by_pose_change"]:
            if inl < 15:
                invalid = True
            else:
                sel = [k for k in range(n) if status[k] == dao.TRACKED]
                pairs = [(int(bf["trainIdx"][k]), k) for k in sel]
                r = self.vo.change_in_pose(pairs, other_kf, this_kf, list(ini_pose))
                for j, k in enumerate(sel):
                    if (not r["valid"]) or r["residual"][j] > o["residual_th"]:
                        status[k] = dao.REJ_POSE
        if invalid:
            return info
        for k in range(n):
            if status[k] == dao.TRACKED:
                info["first"][k] = int(bf["trainIdx"][k]); info["dist"][k] = float(bf["distance"][k]); info["tracked"] += 1
        return info

    # ---- the loop --------------------------------------------------------------------------------------------------------
    def run(self, frames, max_num_kfs=0):
        for f in range(len(frames)):
            left, right = frames[f, 0], frames[f, 1]
            if not self.kfs:
                self._first(left, right)
            elif not self._next(left, right):
                self.count += 1
                continue
            self.frame_poses.append(list(self.cur_pose))
            self.count += 1
            if max_num_kfs > 0 and len(self.kfs) == max_num_kfs:
                break
        return self

    def _first(self, left, right):
        self.vo.process(left, right); self.vo.set_this_frame_as_kf()
        kf = self.vo.get_values(); kf["id"] = self.last_kf_id; self.last_kf_id += 1
        self.kfs.append(kf)
        self.last_match_id = kf["ids"][-1]
        self._bow_insert(kf)
        obs = self._observations(kf)
        self.rba.define_new_keyframe([o[0] for o in obs], [0.0] * 6)
        self.obs_lists.append(obs); self.kf_frames.append(self.count)
        self.last_kf_pose = [0.0] * 6

    def _adaptive(self, left, right):
        o = self.o; vo = self.vo
        repeat = False
        while True:
            res = vo.process(left, right, repeat)
            repeat = False
            m = res["stereo_matches"]
            if o["orb_adaptive_fast_th"]:
                if m < o["adaptive_th_min_matches"]:
                    if not vo.fast_is_min():
                        vo.set_fast(max(0, vo.fast_th - 10)); repeat = True
                    elif not vo.orb_is_max():
                        vo.set_orb(vo.orb_th + 10); repeat = True
                    elif m >= 8:
                        pass
                    else:
                        raise RuntimeError("The number of found matches is less than the minimum. Aborting")
                elif m < o["adaptive_th_min_matches"] * 1.2:
                    if not vo.fast_is_min():
                        vo.set_fast(max(0, vo.fast_th - 5))
                    elif not vo.orb_is_max():
                        vo.set_orb(vo.orb_th + 5)
                else:
                    vo.set_fast(min(o["detect_fast_th"], vo.fast_th + 5)); vo.set_orb(vo.orb_default)
            if not repeat:
                return res

    def _next(self, left, right):
        o = self.o
        res = self._adaptive(left, right)
        if not res["valid"]:
            return False
        inc = res["out_pose"]
        self.cur_pose = geo.rv_compose(self.cur_pose, inc); self.incr_kf = geo.rv_compose(self.incr_kf, inc); self.incr_check = geo.rv_compose(self.incr_check, inc)
        force_new = _norm3(self.incr_kf[3:6]) > self.max_tr_limit or _norm3(self.incr_kf[:3]) > (self.max_rot_limit * math.pi / 180.0)
        force_track = False if o["vo_id_tracking_th"] == 0 else res["tracked_kf"] < o["vo_id_tracking_th"]
        force_dist = _norm3(self.incr_check[3:6]) > self.max_tr or _norm3(self.incr_check[:3]) > (self.max_rot * math.pi / 180.0)
        if force_new or force_track or force_dist:
            self._check(force_new, force_track, force_dist)
        return True

    def _check(self, force_new, force_track, force_dist):
        o = self.o
        rec = dict(frame=self.count, flags=(force_new, force_track, force_dist))
        if force_track:
            self.vo.reset_ids()
        self.incr_check = [0.0] * 6
        kf = self.vo.get_values(); kf["id"] = self.last_kf_id
        self.kfs.append(kf)
        q = orc.bow_query(self.voc, self.db, kf["dl"], 4)
        assert len(q) > 0
        rec["query"] = q
        sim, poses, lc_id, potential = self._similar_kfs(kf["id"], q)
        rec.update(similar=sim, lc_id=lc_id, potential_lc=potential)
        insert = q[0][1] < 0.05
        da = [self._internal_da(kf, self.kfs[s], poses[k] if k < len(poses) else [0.0] * 6) for k, s in enumerate(sim)]
        rec["tracked"] = [(d["kf_idx"], d["tracked"]) for d in da]
        order = sorted(range(len(da)), key=lambda i: -da[i]["tracked"])          # stable
        highest = da[order[0]]["tracked"]
        if force_new:
            insert = True
        confirmed = False
        if potential:
            lc_i = 0
            for i, d in enumerate(da):
                if d["kf_idx"] == lc_id and d["tracked"] > 0.5 * highest:
                    confirmed = True; lc_i = i
                    break
            if confirmed:
                insert = True
                order.remove(lc_i); order.insert(0, lc_i)
        else:
            if highest < o["updated_matches_th"]:
                insert = True
            else:
                olimit = o["updated_matches_th"] + o["up_matches_th_plus"]
                if highest <= olimit:
                    x = highest - o["updated_matches_th"]
                    t = 0.02 + (0.25 / o["up_matches_th_plus"]) * x
                    self.max_tr = min(max(t, 0.02), 0.3)
                    r = 15 + 13 / olimit * (highest - olimit)
                    self.max_rot = min(max(r, 2), 15)
        rec.update(confirmed_lc=confirmed, inserted=insert)
        if not insert:
            self.kfs.pop()
        else:
            ids, self.last_match_id, ntr, nnew = dao.assign_match_ids(len(kf["m"]), [(d["kf_idx"], d["first"]) for d in da], order,
                                                                      [k["ids"] for k in self.kfs], self.last_match_id)
            kf["ids"] = ids; rec.update(n_tracked_ids=ntr, n_new_ids=nnew)
            self.vo.reset_ids()
            self.last_kf_id += 1
            self.max_tr = o["max_translation"]; self.max_rot = o["max_rotation"]
            obs = self._observations(kf)
            init = self._via_cpose3d(self.incr_kf) if o["use_initial_pose"] else [0.0] * 6
            self.rba.define_new_keyframe([x[0] for x in obs], init)
            self.obs_lists.append(obs); self.kf_frames.append(self.count)
            self.last_kf_pose = self._via_cpose3d(self.rba.pose[kf["id"]])
            self._bow_insert(kf)
            self.cur_pose = geo.rv_compose(self.last_kf_pose, self.K)
            self.incr_kf = [0.0] * 6
        self.checks.append(rec)
