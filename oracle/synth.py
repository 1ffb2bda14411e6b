"""TEST INFRASTRUCTURE ONLY -- seeded synthetic inputs shared by oracle/make_golden.py and the tests, so the
golden files only need to store the reference's OUTPUTS."""
import numpy as np


def pckh_inputs(seed=0, B=6, J=16):
    """Random + adversarial heatmaps / label maps / head boxes for the PCKh evaluators."""
    r = np.random.RandomState(seed)
    x = r.randn(B, J, 64, 64).astype(np.float32)
    x[0, 0] = 0                                   # constant map
    x[1, 3, 10, 5] = 9
    x[1, 3, 10, 7] = 9                            # duplicated maximum
    x[2] = x[2].astype(np.float16).astype(np.float32)  # fp16-quantised maps (many ties)
    tgt = np.zeros([B, 64, 64], dtype=np.int64)
    for b in range(B):
        for j in range(J):
            if b != 4 and r.rand() < 0.85:        # image 4 has no annotated joint -> NaN accuracy row
                tgt[b, r.randint(64), r.randint(64)] = j + 1
    rect = r.uniform(0, 64, size=(B, 4)).astype(np.float32)
    z = r.randn(B, J + 1, 64, 64).astype(np.float32)
    e = np.exp(z - z.max(1, keepdims=True))
    x17 = (e / e.sum(1, keepdims=True)).astype(np.float32)
    t14 = r.rand(B, 14, 64, 64).astype(np.float32)
    t14[0, 2] = 0                                 # absent joint
    x14 = (0.7 * t14 + 0.3 * r.rand(B, 14, 64, 64)).astype(np.float32)
    return dict(x=x, target=tgt, rect=rect, x17=x17, t14=t14, x14=x14)


def pckh_near_inputs(seed=0, B=6, J=16):
    """Class-probability maps [B, J+1, 64, 64] whose peaks sit 0..7 px from the labelled joint, with head boxes of
    5..25 px diagonal: both outcomes of every PCKh threshold test occur, including exact-boundary distances
    (integer d2 against 0.3 * diagonal)."""
    r = np.random.RandomState(1000 + seed)
    z = r.randn(B, J + 1, 64, 64).astype(np.float32)
    tgt = np.zeros([B, 64, 64], dtype=np.int64)
    rect = np.zeros([B, 4], dtype=np.float32)
    for b in range(B):
        diag = r.uniform(5, 25)
        if b == 1:
            diag = 10.0 / 0.3  # standard * 0.5 == 5 up to float32 rounding; d2 = 25 sits on the boundary
        ang = r.uniform(0, 2 * np.pi)
        x0, y0 = r.uniform(20, 40, 2)
        rect[b] = (x0, y0, x0 + diag * np.cos(ang), y0 + diag * np.sin(ang))
        for j in range(J):
            if r.rand() < 0.9:
                ly, lx = r.randint(8, 56, 2)
                if tgt[b, ly, lx] != 0:
                    continue
                tgt[b, ly, lx] = j + 1
                oy, ox = ((3, 4), (4, 3), (5, 0), (0, 5))[j % 4] if b == 1 else r.randint(-5, 6, 2)
                z[b, j + 1, ly + oy, lx + ox] = 12.0
    e = np.exp(z - z.max(1, keepdims=True))
    x = (e / e.sum(1, keepdims=True)).astype(np.float32)
    return dict(x17=x, target=tgt, rect=rect)
