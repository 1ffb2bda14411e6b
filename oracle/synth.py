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
