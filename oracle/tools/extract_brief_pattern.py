#!/usr/bin/env python
"""Recover the 256x4 learned rBRIEF sampling pattern (OpenCV's bit_pattern_31_) by
black-box probing of cv2.ORB.compute().  TEST INFRASTRUCTURE -- run once in the build
container (needs cv2); the result is committed as oracle/brief_pattern_31.inc.

Why: the pattern is *data* that lives in OpenCV (an un-vendored dependency of the reference,
SURVEY.md App. A.1); it is not in /root/reference.  Rather than trust memory we read it
out of the cv2 binary's behaviour: with angle=0 test i compares blurred pixels A_i and B_i
(bit = I(A_i) < I(B_i)).  A single bright (dark) impulse at p sets bit i iff p is nearer to
B_i (A_i) than to the other point, inside the 7x7 blur footprint.  From the two response
maps S_i (bright) and T_i (dark) over p in [-R,R]^2 we brute-force the unique (A_i,B_i).
"""
import sys, numpy as np, cv2

R = 19           # probe radius (pattern coords are within [-15,15]; blur adds 3)
C = 100          # keypoint centre in a 201x201 image
BG = 100

def blurred_response(val):
    img = np.full((2*C+1, 2*C+1), BG, np.uint8); img[C, C] = val
    b = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    return b[C-3:C+4, C-3:C+4].astype(np.int32)    # [dy+3, dx+3] value at offset d from impulse

def main(out):
    orb = cv2.ORB_create(nfeatures=10, nlevels=1, edgeThreshold=31, patchSize=31, WTA_K=2)
    kp = [cv2.KeyPoint(float(C), float(C), 31.0, 0.0, 0, 0, -1)]
    n = 2*R+1
    S = np.zeros((256, n, n), bool); T = np.zeros((256, n, n), bool)
    for pol, dst, val in ((0, S, 255), (1, T, 0)):
        for py in range(-R, R+1):
            for px in range(-R, R+1):
                img = np.full((2*C+1, 2*C+1), BG, np.uint8); img[C+py, C+px] = val
                _, d = orb.compute(img, kp)
                dst[:, py+R, px+R] = np.unpackbits(d[0], bitorder='little').astype(bool)
    Rb, Rd = blurred_response(255), blurred_response(0)
    ys, xs = np.mgrid[-R:R+1, -R:R+1]
    def resp(tab, qx, qy):                 # blurred value at q for impulses at every p
        dx, dy = qx - xs, qy - ys
        inside = (abs(dx) <= 3) & (abs(dy) <= 3)
        v = np.full(xs.shape, BG, np.int32)
        v[inside] = tab[dy[inside]+3, dx[inside]+3]
        return v
    cacheb, cached = {}, {}
    def rb(q):
        if q not in cacheb: cacheb[q] = resp(Rb, *q)
        return cacheb[q]
    def rd(q):
        if q not in cached: cached[q] = resp(Rd, *q)
        return cached[q]
    pat = np.zeros((256, 4), np.int8)
    for i in range(256):
        Bc = [(int(xs[m]), int(ys[m])) for m in zip(*np.nonzero(S[i]))]
        Ac = [(int(xs[m]), int(ys[m])) for m in zip(*np.nonzero(T[i]))]
        sols = []
        for B in Bc:
            if abs(B[0]) > 15 or abs(B[1]) > 15: continue
            for A in Ac:
                if abs(A[0]) > 15 or abs(A[1]) > 15: continue
                if np.array_equal(rb(A) < rb(B), S[i]) and np.array_equal(rd(A) < rd(B), T[i]):
                    sols.append((A, B))
        assert len(sols) == 1, (i, len(sols), sols[:4])
        (ax, ay), (bx, by) = sols[0]
        pat[i] = (ax, ay, bx, by)
    with open(out, 'w') as f:
        f.write("/* rBRIEF 31x31 sampling pattern: 256 tests x {ax, ay, bx, by}; bit = I(a) < I(b).\n"
                " * Recovered from cv2 %s by oracle/tools/extract_brief_pattern.py (black-box probing). */\n" % cv2.__version__)
        for i in range(256):
            f.write("%d,%d,%d,%d,%s" % (*pat[i], "\n" if i % 8 == 7 else " "))
    print("wrote", out, "range", pat.min(), pat.max())

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "oracle/brief_pattern_31.inc")
