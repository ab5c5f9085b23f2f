"""oracle/telea_seq.py -- TEST INFRASTRUCTURE ONLY.

Sequential restatement of OpenCV's `cv2.inpaint(img_u8, mask, 1, INPAINT_TELEA)` for one channel
(modules/photo/src/inpaint.cpp: icvTeleaInpaintFMM, icvCalcFMM, FastMarching_solve, CvPriorityQueueFloat).
OpenCV is a third-party dependency of the reference's Inpainting plugin (plugins/inpainting.py:59; `opencv-python`,
unpinned in requirements.txt:5); its source is not under /root/reference, so the algorithm is restated here from its
published description and PINNED against the cv2 4.13 binary in this image (tests/test_oracle_cpu.py).  It documents the
exact event order the device implementation (csrc, k_inpaint_*) reproduces.
"""
import numpy as np

KNOWN, BAND, INSIDE = 0, 1, 2


def _solve(f, t, i1, j1, i2, j2):
    a11 = float(t[i1, j1]); a22 = float(t[i2, j2])
    m12 = min(a11, a22)
    if f[i1, j1] != INSIDE:
        if f[i2, j2] != INSIDE:
            if abs(a11 - a22) >= 1.0:
                sol = 1 + m12
            else:
                sol = (a11 + a22 + np.sqrt(2 - (a11 - a22) * (a11 - a22))) * 0.5
        else:
            sol = 1 + a11
    elif f[i2, j2] != INSIDE:
        sol = 1 + a22
    else:
        sol = 1 + m12
    return np.float32(sol)


class _Heap:
    """Sorted list with FIFO order among equal keys (CvPriorityQueueFloat: Push walks back while prev.T > T)."""

    def __init__(self):
        import bisect
        self.bisect = bisect
        self.keys = []
        self.items = []

    def push(self, i, j, T):
        k = self.bisect.bisect_right(self.keys, float(T))
        self.keys.insert(k, float(T)); self.items.insert(k, (i, j))

    def pop(self):
        if not self.keys:
            return None
        self.keys.pop(0)
        return self.items.pop(0)


def _calc_fmm(f, t, heap, negate):
    """icvCalcFMM: distance of the pixels flagged INSIDE in f from the band in the heap"""
    rows, cols = f.shape
    while True:
        it = heap.pop()
        if it is None:
            break
        ii, jj = it
        f[ii, jj] = KNOWN
        for q in range(4):
            i, j = [(ii - 1, jj), (ii, jj - 1), (ii + 1, jj), (ii, jj + 1)][q]
            if i <= 0 or j <= 0 or i > rows - 1 or j > cols - 1:
                continue
            if i > rows - 2 or j > cols - 2:     # solve reads i+1 / j+1
                continue
            if f[i, j] == INSIDE:
                dist = min(_solve(f, t, i - 1, j, i, j - 1), _solve(f, t, i + 1, j, i, j - 1),
                           _solve(f, t, i - 1, j, i, j + 1), _solve(f, t, i + 1, j, i, j + 1))
                t[i, j] = dist
                f[i, j] = BAND
                heap.push(i, j, dist)
    if negate:
        for i in range(rows):
            for j in range(cols):
                if f[i, j] == KNOWN and True:
                    pass
    return


def telea_radius1(img, mask, return_state=False):
    """img (H,W) uint8, mask (H,W) nonzero = to inpaint.  Returns the inpainted uint8 image."""
    H, W = img.shape
    rows, cols = H + 2, W + 2
    rng = 1
    f = np.zeros((rows, cols), np.uint8)
    t = np.full((rows, cols), 1.0e6, np.float32)
    m = np.zeros((rows, cols), np.uint8)
    m[1:-1, 1:-1] = (mask != 0)
    out = img.copy()
    # band = dilate(mask, cross) - mask
    dil = m.copy()
    dil[1:, :] |= m[:-1, :]; dil[:-1, :] |= m[1:, :]; dil[:, 1:] |= m[:, :-1]; dil[:, :-1] |= m[:, 1:]
    band = dil & (1 - m)
    band[0, :] = 0; band[-1, :] = 0; band[:, 0] = 0; band[:, -1] = 0
    heap = _Heap()
    for i in range(rows):
        for j in range(cols):
            if band[i, j]:
                heap.push(i, j, 0.0)
    f[band != 0] = BAND
    f[m != 0] = INSIDE
    t[band != 0] = 0
    # TELEA: signed distance outside, within `range` of the mask
    el = np.zeros((rows, cols), np.uint8)
    el |= m
    el[1:, :] |= m[:-1, :]; el[:-1, :] |= m[1:, :]; el[:, 1:] |= m[:, :-1]; el[:, :-1] |= m[:, 1:]   # ellipse 3x3 == cross
    outm = el & (1 - m)                      # dilate(mask, el_range) - mask
    outm2 = outm & (1 - band)                # cvSub(out, band, out)
    outm2[0, :] = 0; outm2[-1, :] = 0; outm2[:, 0] = 0; outm2[:, -1] = 0
    # with range 1 the outside set minus the band is empty: nothing to march; t outside stays as is
    # ---- main FMM
    order = []
    while True:
        it = heap.pop()
        if it is None:
            break
        ii, jj = it
        f[ii, jj] = KNOWN
        for q in range(4):
            i, j = [(ii - 1, jj), (ii, jj - 1), (ii + 1, jj), (ii, jj + 1)][q]
            if i <= 1 - 1 or j <= 1 - 1 or i > rows - 1 or j > cols - 1:
                continue
            if i <= 0 or j <= 0 or i >= rows - 1 or j >= cols - 1:
                continue
            if f[i, j] != INSIDE:
                continue
            dist = min(_solve(f, t, i - 1, j, i, j - 1), _solve(f, t, i + 1, j, i, j - 1),
                       _solve(f, t, i - 1, j, i, j + 1), _solve(f, t, i + 1, j, i, j + 1))
            t[i, j] = dist
            # gradT
            if f[i, j + 1] != INSIDE:
                if f[i, j - 1] != INSIDE:
                    gtx = np.float32((t[i, j + 1] - t[i, j - 1])) * np.float32(0.5)
                else:
                    gtx = np.float32(t[i, j + 1] - t[i, j])
            else:
                gtx = np.float32(t[i, j] - t[i, j - 1]) if f[i, j - 1] != INSIDE else np.float32(0)
            if f[i + 1, j] != INSIDE:
                if f[i - 1, j] != INSIDE:
                    gty = np.float32((t[i + 1, j] - t[i - 1, j])) * np.float32(0.5)
                else:
                    gty = np.float32(t[i + 1, j] - t[i, j])
            else:
                gty = np.float32(t[i, j] - t[i - 1, j]) if f[i - 1, j] != INSIDE else np.float32(0)
            Ia = np.float32(0); Jx = np.float32(0); Jy = np.float32(0); s = np.float32(1.0e-20)
            for k in range(i - rng, i + rng + 1):
                km = k - 1 + (k == 1); kp = k - 1 - (k == rows - 2)
                for l in range(j - rng, j + rng + 1):
                    lm = l - 1 + (l == 1); lp = l - 1 - (l == cols - 2)
                    if k > 0 and l > 0 and k < rows - 1 and l < cols - 1:
                        if f[k, l] != INSIDE and ((l - j) * (l - j) + (k - i) * (k - i) <= rng * rng):
                            ry = np.float32(i - k); rx = np.float32(j - l)
                            vl = np.float32(rx * rx + ry * ry)
                            dst = np.float32(1.0 / (float(vl) * np.sqrt(float(vl))))
                            lev = np.float32(1.0 / (1 + abs(float(t[k, l]) - float(t[i, j]))))
                            dirv = np.float32(rx * gtx + ry * gty)
                            if abs(dirv) <= 0.01:
                                dirv = np.float32(0.000001)
                            w = np.float32(abs(np.float32(np.float32(dst * lev) * dirv)))
                            if f[k, l + 1] != INSIDE:
                                if f[k, l - 1] != INSIDE:
                                    gix = np.float32(int(out[km, lp + 1]) - int(out[km, lm - 1])) * np.float32(2.0)
                                else:
                                    gix = np.float32(int(out[km, lp + 1]) - int(out[km, lm]))
                            else:
                                gix = np.float32(int(out[km, lp]) - int(out[km, lm - 1])) if f[k, l - 1] != INSIDE else np.float32(0)
                            if f[k + 1, l] != INSIDE:
                                if f[k - 1, l] != INSIDE:
                                    giy = np.float32(int(out[kp + 1, lm]) - int(out[km - 1, lm])) * np.float32(2.0)
                                else:
                                    giy = np.float32(int(out[kp + 1, lm]) - int(out[km, lm]))
                            else:
                                giy = np.float32(int(out[kp, lm]) - int(out[km - 1, lm])) if f[k - 1, l] != INSIDE else np.float32(0)
                            Ia = np.float32(Ia + w * np.float32(out[k - 1, l - 1]))
                            Jx = np.float32(Jx - w * np.float32(gix * rx))
                            Jy = np.float32(Jy - w * np.float32(giy * ry))
                            s = np.float32(s + w)
            # all in float: sqrt resolves to the float overload, 1.0e-20f and 0.5f are float literals
            den = np.float32(np.float32(np.sqrt(np.float32(Jx * Jx + Jy * Jy))) + np.float32(1.0e-20))
            sat = np.float32(np.float32(np.float32(Ia / s) + np.float32(np.float32(Jx + Jy) / den)) + np.float32(0.5))
            out[i - 1, j - 1] = np.uint8(min(max(int(np.rint(sat)), 0), 255))
            f[i, j] = BAND
            heap.push(i, j, dist)
            order.append((i, j))
    if return_state:
        return out, t, order
    return out
