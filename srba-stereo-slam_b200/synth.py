"""Synthetic KITTI-shaped stereo sequences rendered from a known 3-D scene.

The reference ships no images (demo/config_imgdir_kitti_srba.ini:91 points at an absent `dataset0`), so
tests and the benchmark render a seeded corridor scene: a textured ground plane, two side walls, a far
wall and textured boxes, seen by a rectified pinhole stereo rig with the shipped intrinsics
(fx=fy=707.0912, cx=601.8873, cy=183.1104, baseline 0.54 m; kitti.ini:138-155) moving ~0.8 m per
frame along a gently curving closed loop.  Ground-truth poses are returned with the frames.
Pure numpy; deterministic for a given seed.
"""
import numpy as np

FX = FY = 707.0912
CX, CY = 601.8873, 183.1104
BASELINE = 0.54


def _texture_pyramid(rng, size=1024, levels=7):
    """Multi-scale random texture in [0,1] with blocky structure (corners) and smooth shading."""
    t = np.zeros((size, size), np.float32)
    for cell, amp in ((4, 0.10), (8, 0.18), (16, 0.22), (32, 0.22), (64, 0.18), (128, 0.10)):
        n = rng.random((size // cell, size // cell)).astype(np.float32)
        t += amp * np.kron(n, np.ones((cell, cell), np.float32))
    t = (t - t.min()) / (t.max() - t.min())
    pyr = [t]
    for _ in range(levels - 1):
        p = pyr[-1]
        pyr.append(0.25 * (p[0::2, 0::2] + p[1::2, 0::2] + p[0::2, 1::2] + p[1::2, 1::2]))
    return pyr


def _sample(pyr, u, v, footprint, texel):
    """Bilinear lookup of world-plane coordinates (u, v) [m] at the mip level matching the pixel footprint."""
    lvl = np.clip(np.floor(np.log2(np.maximum(footprint / texel, 1.0))).astype(np.int32), 0, len(pyr) - 1)
    out = np.zeros(u.shape, np.float32)
    for l in np.unique(lvl):
        m = lvl == l
        p = pyr[l]; n = p.shape[0]
        s = 1.0 / (texel * (1 << l))
        x = u[m] * s; y = v[m] * s
        x0 = np.floor(x).astype(np.int64); y0 = np.floor(y).astype(np.int64)
        fx = (x - x0).astype(np.float32); fy = (y - y0).astype(np.float32)
        x0 %= n; y0 %= n; x1 = (x0 + 1) % n; y1 = (y0 + 1) % n
        out[m] = (p[y0, x0] * (1 - fx) * (1 - fy) + p[y0, x1] * fx * (1 - fy) +
                  p[y1, x0] * (1 - fx) * fy + p[y1, x1] * fx * fy)
    return out


class CorridorScene:
    def __init__(self, seed=0, width=1241, height=376, fx=FX, fy=FY, cx=CX, cy=CY, baseline=BASELINE,
                 loop_radius=60.0, step=0.8, noise_sigma=2.0):
        self.seed = seed; self.W = width; self.H = height
        self.fx, self.fy, self.cx, self.cy, self.b = fx * width / 1241.0, fy * width / 1241.0, cx * width / 1241.0, cy * height / 376.0, baseline
        self.R = loop_radius; self.step = step; self.sigma = noise_sigma
        rng = np.random.default_rng(seed)
        self.tex = [_texture_pyramid(rng) for _ in range(4)]          # ground, inner wall, outer wall, boxes
        self.texel = 0.02
        self.cam_h = 1.65; self.half_w = 6.0; self.wall_h = 8.0
        nb = 80                                                       # boxes: billboards along the ring
        self.box_ang = rng.uniform(0, 2 * np.pi, nb)
        self.box_rad = self.R + rng.uniform(-4.5, 4.5, nb)
        self.box_size = rng.uniform(0.8, 2.5, (nb, 2))
        self.box_shade = rng.uniform(0.3, 1.0, nb)
        u, v = np.meshgrid(np.arange(self.W, dtype=np.float64), np.arange(self.H, dtype=np.float64))
        self.rays = np.stack([(u - self.cx) / self.fx, (v - self.cy) / self.fy, np.ones_like(u)], -1)   # camera: x right, y down, z fwd

    def pose(self, k):
        """Ground-truth left-camera pose of frame k: position on a ring of radius R, heading tangent.
        Returns (R_wc [3x3], t_w [3]); world: X east, Y down, Z north."""
        th = k * self.step / self.R
        pos = np.array([self.R * np.cos(th), 0.0, self.R * np.sin(th)])
        fwd = np.array([-np.sin(th), 0.0, np.cos(th)])
        right = np.array([np.cos(th), 0.0, np.sin(th)])     # radially outward is to the right of travel (ccw ring seen from above with Y down)
        down = np.array([0.0, 1.0, 0.0])
        Rwc = np.stack([right, down, fwd], 1)
        return Rwc, pos

    def _render(self, Rwc, pos, rng):
        d = self.rays @ Rwc.T                       # world ray directions, (H, W, 3)
        ox, oy, oz = pos
        dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
        best_t = np.full(dx.shape, 1e9); img = np.full(dx.shape, 0.75, np.float32)    # sky
        img -= 0.25 * (np.arange(self.H, dtype=np.float32)[:, None] / self.H)
        fp_scale = 1.0 / self.fx
        # ground plane y = cam_h
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (self.cam_h - oy) / dy
        m = (dy > 1e-9) & (t > 0.1) & (t < best_t)
        if m.any():
            X = ox + t[m] * dx[m]; Z = oz + t[m] * dz[m]
            rad = np.hypot(X, Z)
            ok = np.abs(rad - self.R) < self.half_w
            val = _sample(self.tex[0], X, Z, t[m] * fp_scale * np.maximum(1.0, 1.0 / np.maximum(dy[m], 0.05) * 0.3), self.texel)
            val = np.where(ok, 0.25 + 0.5 * val, 0.2 + 0.25 * val)
            img[m] = val; best_t[m] = t[m]
        # cylindrical walls at radius R -/+ half_w (ray-cylinder intersection in the XZ plane)
        a = dx * dx + dz * dz
        bq = 2 * (ox * dx + oz * dz)
        for wi, rr in enumerate((self.R - self.half_w, self.R + self.half_w)):
            cq = ox * ox + oz * oz - rr * rr
            disc = bq * bq - 4 * a * cq
            with np.errstate(invalid="ignore", divide="ignore"):
                sq = np.sqrt(np.maximum(disc, 0))
                t1 = (-bq - sq) / (2 * a); t2 = (-bq + sq) / (2 * a)
            t = np.where((t1 > 0.1) & (wi == 0), t1, t2)       # inner wall: first hit; outer wall: far root
            y = oy + t * dy
            m = (disc > 0) & (t > 0.1) & (t < best_t) & (y < self.cam_h) & (y > self.cam_h - self.wall_h)
            if m.any():
                X = ox + t[m] * dx[m]; Z = oz + t[m] * dz[m]
                ang = np.arctan2(Z, X) * rr
                val = _sample(self.tex[1 + wi], ang, y[m] + 20.0, t[m] * fp_scale, self.texel)
                img[m] = 0.15 + 0.7 * val; best_t[m] = t[m]
        # boxes (vertical billboards facing the ring tangent)
        for k in range(len(self.box_ang)):
            c = np.array([self.box_rad[k] * np.cos(self.box_ang[k]), self.cam_h - self.box_size[k, 1] / 2, self.box_rad[k] * np.sin(self.box_ang[k])])
            if np.hypot(c[0] - ox, c[2] - oz) > 45.0:
                continue
            nrm = np.array([-np.sin(self.box_ang[k]), 0.0, np.cos(self.box_ang[k])])
            den = dx * nrm[0] + dz * nrm[2]
            with np.errstate(divide="ignore", invalid="ignore"):
                t = ((c[0] - ox) * nrm[0] + (c[2] - oz) * nrm[2]) / den
            X = ox + t * dx; Y = oy + t * dy; Z = oz + t * dz
            lu = (X - c[0]) * np.cos(self.box_ang[k]) + (Z - c[2]) * np.sin(self.box_ang[k])
            lv = Y - c[1]
            m = (np.abs(den) > 1e-9) & (t > 0.1) & (t < best_t) & (np.abs(lu) < self.box_size[k, 0] / 2) & (np.abs(lv) < self.box_size[k, 1] / 2)
            if m.any():
                val = _sample(self.tex[3], lu[m] + 7.0 * k, lv[m] + 3.0 * k, t[m] * fp_scale, self.texel)
                img[m] = self.box_shade[k] * (0.2 + 0.8 * val); best_t[m] = t[m]
        out = 255.0 * img + rng.normal(0.0, self.sigma, img.shape)
        return np.clip(np.rint(out), 0, 255).astype(np.uint8)

    def frame(self, k):
        """Rectified stereo pair (left, right) of frame k, uint8 [H, W] each."""
        Rwc, pos = self.pose(k)
        rng = np.random.default_rng([self.seed, k])
        left = self._render(Rwc, pos, rng)
        right = self._render(Rwc, pos + Rwc[:, 0] * self.b, rng)
        return left, right

    def sequence(self, n, start=0):
        out = np.zeros((n, 2, self.H, self.W), np.uint8)
        for i in range(n):
            out[i, 0], out[i, 1] = self.frame(start + i)
        return out


def stereo_batch(npairs, width=1241, height=376, seed=0, unique=None, augment=True):
    """[npairs, 2, H, W] uint8.  `unique` frames (default min(npairs, 8)) are rendered; with `augment` each is also
    mirrored horizontally (left/right swapped so disparities stay positive), flipped vertically and intensity-inverted,
    giving 8 distinct valid rectified stereo pairs per rendered frame (different descriptors and BoW words) before the
    set is tiled to npairs.  Rendering 1241x376 costs ~1.6 s per image, hence the reuse."""
    unique = min(npairs, 8) if unique is None else min(unique, npairs)
    sc = CorridorScene(seed=seed, width=width, height=height)
    base = sc.sequence(unique)
    if augment:
        var = []
        for inv in (False, True):
            for vf in (False, True):
                for hf in (False, True):
                    v = base
                    if hf:
                        v = v[:, ::-1, :, ::-1]            # mirror columns and swap the two views
                    if vf:
                        v = v[:, :, ::-1, :]
                    if inv:
                        v = 255 - v
                    var.append(v)
        base = np.stack(var, 1).reshape(-1, 2, height, width)   # frame-major: consecutive pairs are different scenes
    reps = (npairs + len(base) - 1) // len(base)
    return np.ascontiguousarray(np.tile(base, (reps, 1, 1, 1))[:npairs])


def _render_job(args):
    seed, W, H, radius, k = args
    return CorridorScene(seed=seed, width=W, height=H, loop_radius=radius).frame(k)


def render_sequence_parallel(n, width, height, loop_radius, seed, workers=None):
    """[n, 2, H, W] uint8, frames rendered by a pool of worker processes (1241x376 costs ~1.2 s per image on one core)."""
    import os
    from concurrent.futures import ProcessPoolExecutor
    jobs = [(seed, width, height, loop_radius, k) for k in range(n)]
    workers = workers or max(1, min(len(os.sched_getaffinity(0)), 32))
    if workers == 1 or n < 8:
        fr = [_render_job(j) for j in jobs]
    else:
        with ProcessPoolExecutor(workers) as ex:
            fr = list(ex.map(_render_job, jobs, chunksize=2))
    out = np.zeros((n, 2, height, width), np.uint8)
    for i, (l, r) in enumerate(fr):
        out[i, 0], out[i, 1] = l, r
    return out


if __name__ == "__main__":
    # python synth.py <n> <W> <H> <loop_radius> <seed> <out.bin>: renders from a fresh single-threaded process (safe to fork workers from)
    import sys
    n, W, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    render_sequence_parallel(n, W, H, float(sys.argv[4]), int(sys.argv[5])).tofile(sys.argv[6])
