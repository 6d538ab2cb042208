#!/usr/bin/env python
"""numpy model of pyr_down_roll_kernel's index and border logic (slam-sdvl_b200/csrc/pyramid.cu), checked against the
direct 25-tap definition of cv::pyrDown on small images of every row alignment (16 / 8 / 4 bytes), partial row blocks,
partial column groups and all-255 images (the 16-bit lanes must not carry).  It was run before the kernel's first GPU
test; the GPU parity tests (tests/test_gpu_parity.py::test_pyramid_*) are what pins the kernel itself.

  python profiles/scripts/pyr_roll_model.py"""
import numpy as np


def refl(t, n):
    t = max(-(n - 1), min(2 * n - 2, t))
    return -t if t < 0 else (2 * n - 2 - t if t >= n else t)


def direct(src):
    sh, sw = src.shape
    dh, dw = sh // 2, sw // 2
    k = np.array([1, 4, 6, 4, 1])
    out = np.zeros((dh, dw), np.uint8)
    for y in range(dh):
        rows = [refl(2 * y + i - 2, sh) for i in range(5)]
        for x in range(dw):
            cols = [refl(2 * x + j - 2, sw) for j in range(5)]
            out[y, x] = (int((k[:, None] * k[None, :] * src[np.ix_(rows, cols)].astype(np.int64)).sum()) + 128) >> 8
    return out


def dp4a(w, c, acc):
    return acc + sum(((w >> (8 * b)) & 255) * ((c >> (8 * b)) & 255) for b in range(4))


def word(row, c):
    return int(row[c]) | int(row[c + 1]) << 8 | int(row[c + 2]) << 16 | int(row[c + 3]) << 24


def load_row8(row, cm, sw, align):          # w[k] = columns cm - 4 + 4k .. + 3; words outside the row read as 0
    w = [0] * 6
    w[0] = word(row, cm - 4) if cm > 0 else 0
    if align == 16:
        for k in range(1, 5):
            w[k] = word(row, cm + 4 * k - 4)
    elif align == 8:
        w[1], w[2] = word(row, cm), word(row, cm + 4)
        if cm + 16 <= sw:
            w[3], w[4] = word(row, cm + 8), word(row, cm + 12)
    else:
        for k in range(1, 5):
            w[k] = word(row, cm + 4 * k - 4) if cm + 4 * k <= sw else 0
    w[5] = word(row, cm + 16) if cm + 20 <= sw else 0
    return w


def hsum8(w, coef_e0, coef_o):              # 8 horizontal sums as four 16-bit pairs
    h = []
    for j in range(4):
        e = dp4a(w[j + 1], coef_e0 if j == 0 else 0x00010406, dp4a(w[j], 0x04010000, 0))
        o = dp4a(w[j + 2], 1, dp4a(w[j + 1], coef_o[j], 0))
        h.append((e + (o << 16)) & 0xFFFFFFFF)
    return h


def rolling(src, PR):
    sh, sw = src.shape
    dh, dw = sh // 2, sw // 2
    align = 16 if sw % 16 == 0 else 8 if sw % 8 == 0 else 4
    out = np.zeros((dh, dw), np.uint8)
    row = lambda s: src[refl(min(s, sh), sh)]
    for rb in range((dh + PR - 1) // PR):
        for g in range((dw + 7) >> 3):
            y0, x0 = rb * PR, g * 8
            cm = 2 * x0
            coef_e0 = 0x00020806 if cm == 0 else 0x00010406                 # columns -2, -1 <- 2, 1
            coef_o = [0x04070401 if x0 + 2 * j + 1 == dw - 1 else 0x04060401 for j in range(4)]   # column sw <- sw - 2
            H = lambda s: hsum8(load_row8(row(s), cm, sw, align), coef_e0, coef_o)
            ha, hb, hc = H(2 * y0 - 2), H(2 * y0 - 1), H(2 * y0)
            for r in range(PR):
                y = y0 + r
                hd, he = H(2 * y + 1), H(2 * y + 2)
                v = [(ha[j] + he[j] + 0x00800080 + 4 * (hb[j] + hd[j]) + 6 * hc[j]) & 0xFFFFFFFF for j in range(4)]
                o8 = [b for j in range(4) for b in ((v[j] >> 8) & 255, (v[j] >> 24) & 255)]
                if y < dh:
                    for k in range(8):
                        if x0 + k < dw:
                            out[y, x0 + k] = o8[k]
                ha, hb, hc = hc, hd, he
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    ok = True
    for (w, h) in [(48, 20), (40, 22), (44, 18), (36, 14), (16, 8), (188, 30), (20, 9), (12, 7), (32, 34), (8, 6), (4, 4)]:
        for kind in range(2):
            for PR in (4, 8):
                src = rng.integers(0, 256, (h, w), dtype=np.uint8) if kind == 0 else np.full((h, w), 255, np.uint8)
                same = np.array_equal(direct(src), rolling(src, PR))
                ok &= same
                if not same:
                    print("MISMATCH", w, h, "random" if kind == 0 else "all-255", "rows/thread", PR)
    print("rolling form == direct form on every case" if ok else "FAILED")
    raise SystemExit(0 if ok else 1)
