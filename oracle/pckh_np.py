"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's heatmap decoding and PCKh evaluators.

Integer results (decoded indices, correct / total counts) are the bit-exact parity target of the CUDA kernels.
Float decisions replicate the reference's dtypes: distances and `standard` are float32 torch scalars and the
float64 thresholds of np.arange(0, 0.55, 0.05) are rounded to float32 by the tensor-vs-scalar comparison
(SURVEY Appendix D).  Parity pin: tests/test_oracle_pckh.py runs these against the real PCKh classes of
/root/reference on random and adversarial heatmaps (ties, constant maps, absent joints).
"""
import numpy as np

THRESHOLDS_F32 = np.arange(0, 0.55, 0.05).astype(np.float32)  # hourglass_compare.py:835


def argmax_first(hm):
    """First row-major (y, x) of the maximum: torch.nonzero(x >= x.max())[0] (hourglass_compare.py:831) and
    the row-max-then-col form of only_one_hourgless.py:294-295 agree on this."""
    hm = np.asarray(hm)
    idx = int(np.argmax(hm.reshape(-1)))  # numpy argmax returns the first maximal index
    return idx // hm.shape[1], idx % hm.shape[1]


def pckh_sweep(x, target, rect, chan_offset=0, njoints=None):
    """PCKh 'C' (chan_offset=0: hourglass_compare.py:812-844, performance_compare.py:581-615) and
    PCKh 'B' (chan_offset=1, njoints=C-1: performance_compare.py:544-578, train.py:759-791).

    x [B,C,H,W] float, target [B,H,W] int64, rect [B,4] float32.
    Returns dict(correct[B,11] int, total[B,11] int, predict[B,nj,2] (x,y), label[B,nj,2], found[B,nj], standard[B],
    accuracy[B,11] float64 (NaN where total == 0)).
    """
    x = np.asarray(x)
    target = np.asarray(target)
    rect = np.asarray(rect, dtype=np.float32)
    B, C = x.shape[:2]
    nj = C - chan_offset if njoints is None else njoints
    nthr = len(THRESHOLDS_F32)
    correct = np.zeros([B, nthr], dtype=np.int32)
    total = np.zeros([B, nthr], dtype=np.int32)
    predict = np.zeros([B, nj, 2], dtype=np.int32)
    label = np.zeros([B, nj, 2], dtype=np.int32)
    found = np.zeros([B, nj], dtype=np.int32)
    standard = np.zeros([B], dtype=np.float32)
    for i in range(B):
        dx = np.float32(rect[i, 0] - rect[i, 2])
        dy = np.float32(rect[i, 1] - rect[i, 3])
        std = np.float32(np.sqrt(np.float32(np.float32(dx * dx) + np.float32(dy * dy)))) * np.float32(0.6)
        standard[i] = std
        for j in range(nj):
            pos = np.argwhere(target[i] == (j + 1))
            if pos.shape[0] == 0:
                continue
            ly, lx = int(pos[0, 0]), int(pos[0, 1])
            py, px = argmax_first(x[i, j + chan_offset].astype(np.float32))
            d2 = (ly - py) ** 2 + (lx - px) ** 2
            dist = np.float32(np.sqrt(np.float32(d2))) / np.float32(std)
            for s, k in enumerate(THRESHOLDS_F32):
                if np.float32(dist) < k:
                    correct[i, s] += 1
                total[i, s] += 1
            predict[i, j] = (px, py)
            label[i, j] = (lx, ly)
            found[i, j] = 1
    with np.errstate(divide="ignore", invalid="ignore"):
        accuracy = correct.astype(np.float64) / total.astype(np.float64)
    return dict(correct=correct, total=total, predict=predict, label=label, found=found, standard=standard,
                accuracy=accuracy)


def pckh_d(x, target, rect):
    """PCKh 'D' (calculate_parameters.py:906-937): prediction channel j+1, single test
    sqrt(float32(d2)) < standard * 0.5 with standard = sqrt(...) * 0.6, all float32.
    Returns (correct[B], total[B], predict[B,C,2] (x,y), label[B,C,2]); rows of absent joints stay zero."""
    x = np.asarray(x)
    target = np.asarray(target)
    rect = np.asarray(rect, dtype=np.float32)
    B, C = x.shape[:2]
    correct = np.zeros([B], dtype=np.int64)
    total = np.zeros([B], dtype=np.int64)
    predict = np.zeros([B, C, 2], dtype=np.float64)
    label = np.zeros([B, C, 2], dtype=np.float64)
    for i in range(B):
        dx = np.float32(rect[i, 0] - rect[i, 2])
        dy = np.float32(rect[i, 1] - rect[i, 3])
        std = np.float32(np.sqrt(np.float32(np.float32(dx * dx) + np.float32(dy * dy)))) * np.float32(0.6)
        for j in range(C - 1):
            pos = np.argwhere(target[i] == (j + 1))
            if pos.shape[0] == 0:
                continue
            ly, lx = int(pos[0, 0]), int(pos[0, 1])
            py, px = argmax_first(x[i, j + 1].astype(np.float32))
            d2 = (ly - py) ** 2 + (lx - px) ** 2
            if np.float32(np.sqrt(np.float32(d2))) < np.float32(np.float32(std) * np.float32(0.5)):
                correct[i] += 1
            total[i] += 1
            predict[i, j] = (px, py)
            label[i, j] = (lx, ly)
    return correct, total, predict, label


def pckh_a(x, target, batch_size, njoints=14, head_ch=13, neck_ch=1):
    """PCKh 'A' (only_one_hourgless.py:285-313 = try_with_torch_100.py:283-311), quirk Q8 kept: label_xs and
    predict_xs are both the arg-max of the LABEL map's row `head_ys`, so only the y error counts.
    Returns (correct, total)."""
    x = np.asarray(x, dtype=np.float32)
    target = np.asarray(target, dtype=np.float32)
    correct = 0
    total = 0
    for i in range(batch_size):
        hy, hx = argmax_first(target[i, head_ch])
        ny, nx = argmax_first(target[i, neck_ch])
        standard = np.float32(np.sqrt(np.float32((hy - ny) ** 2 + (hx - nx) ** 2))) / np.float32(2)
        for j in range(njoints):
            lab = target[i, j]
            if lab.max() == 0:
                continue
            ly, _ = argmax_first(lab)
            lxs = int(np.argmax(lab[hy]))
            py, _ = argmax_first(x[i, j])
            pxs = lxs
            if np.float32(np.sqrt(np.float32((ly - py) ** 2 + (lxs - pxs) ** 2))) < standard:
                correct += 1
            total += 1
    return correct, total
