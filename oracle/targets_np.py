"""TEST INFRASTRUCTURE ONLY -- numpy / pure-python restatement of the reference's target rendering.

Gaussian keypoint heatmaps follow the numpy expressions of the reference datasets (float64 arithmetic, float32
at the very end through torch.Tensor(...)); the label-map rasteriser restates the point / line primitives of
Pillow's ImageDraw (third-party dependency of the reference, unpinned there; 12.2.0 in this image; C source
src/libImaging/Draw.c: point8, line8, and ImageDraw.line's explicit final point).
Parity pin: tests/test_oracle_targets.py checks these against Pillow itself on random draw lists and against
the reference's own `myImageDataset_COCO.__getitem__` (through a fake COCO shim) when /root/reference exists.
"""
import numpy as np


def centres(kp, size, grid, center_mode=0, truncate=False):
    """kp / w * 64 (try_with_torch.py:110-111) or kp * 256 / w / 4 (hourglass_compare.py:717-718)."""
    kp = np.asarray(kp, dtype=np.float64)
    c = kp / size * grid if center_mode == 0 else kp * 256.0 / size / 4.0
    if truncate:
        c = np.trunc(c)  # .astype(np.int) truncates toward zero
    return c


def gauss_map(persons, img_wh, J, H=64, W=64, center_mode=0, truncate=True, accumulate=False, pre_scale=1.0,
              sigma=1.0, amplitude=1.0):
    """One image.  persons: array [P, J, 3] of (x, y, v).  Returns float32 [J, H, W].

    accumulate=False reproduces quirk Q7 of try_with_torch.py:107-132 / try_different_stack.py:121-144: the map is
    re-zeroed for every person, so only the LAST person's visible joints survive.  accumulate=True is the `+=`
    form of hourglass_compare.py:286-313,713-734.
    """
    persons = np.asarray(persons, dtype=np.float64).reshape(-1, J, 3)
    g = np.zeros([J, H, W])
    xs = np.arange(W, dtype=np.float64)[None, :].repeat(H, 0)
    ys = np.arange(H, dtype=np.float64)[:, None].repeat(W, 1)
    for p in range(persons.shape[0]):
        if not accumulate:
            g = np.zeros([J, H, W])
        for k in range(J):
            if persons[p, k, 2] > 0:
                cx = centres(persons[p, k, 0], img_wh[0], W, center_mode, truncate)
                cy = centres(persons[p, k, 1], img_wh[1], H, center_mode, truncate)
                t = (xs - cx) ** 2 + (ys - cy) ** 2
                if pre_scale != 1.0:
                    t = pre_scale * t
                t = t / (2 * sigma ** 2)
                e = np.exp(-t) if amplitude == 1.0 else amplitude * np.exp(-t)
                if accumulate:
                    g[k] += e
                else:
                    g[k] = e
    return g.astype(np.float32)


def _put(canvas, x, y, ink):
    h, w = canvas.shape
    if 0 <= x < w and 0 <= y < h:
        canvas[y, x] = ink


def draw_point(canvas, x, y, ink):
    """ImageDraw.point on mode 'L': coordinates truncated toward zero, clipped."""
    _put(canvas, int(x), int(y), ink)


def draw_line(canvas, x0, y0, x1, y1, ink):
    """ImageDraw.line (width 0) on mode 'L': Pillow's line8 Bresenham plus the explicit last point."""
    x0, y0, x1, y1 = int(x0), int(y0), int(x1), int(y1)
    ex, ey = x1, y1
    dx, dy = x1 - x0, y1 - y0
    xs = -1 if dx < 0 else 1
    ys = -1 if dy < 0 else 1
    dx, dy = abs(dx), abs(dy)
    if dx == 0:
        for _ in range(dy):
            _put(canvas, x0, y0, ink)
            y0 += ys
    elif dy == 0:
        for _ in range(dx):
            _put(canvas, x0, y0, ink)
            x0 += xs
    elif dx > dy:
        n = dx
        dy += dy
        e = dy - dx
        dx += dx
        for _ in range(n):
            _put(canvas, x0, y0, ink)
            if e >= 0:
                y0 += ys
                e -= dx
            e += dy
            x0 += xs
    else:
        n = dy
        dx += dx
        e = dx - dy
        dy += dy
        for _ in range(n):
            _put(canvas, x0, y0, ink)
            if e >= 0:
                x0 += xs
                e -= dy
            e += dx
            y0 += ys
    _put(canvas, ex, ey, ink)


def label_map(persons, img_wh, J, limbs, H=64, W=64, center_mode=0, draw_points=False, draw_lines=True,
              line_value=0):
    """One image -> int64 [H, W] label map (try_different_stack.py:114-155; try_skeleton_and_keypoints.py:93-114).
    Points carry value k+1, limbs value i+1 (or `line_value` when positive: the background map uses 1; negative: the
    limb index i itself, try_skeleton_from_keypoints_merge.py:130-133)."""
    persons = np.asarray(persons, dtype=np.float64).reshape(-1, J, 3)
    canvas = np.zeros([H, W], dtype=np.uint8)
    for p in range(persons.shape[0]):
        x = centres(persons[p, :, 0], img_wh[0], W, center_mode, True)
        y = centres(persons[p, :, 1], img_wh[1], H, center_mode, True)
        v = persons[p, :, 2]
        if draw_points == 2:
            # MPII keypoint map (train.py:681-686): ellipse((x-.5, y-.5, x+.5, y+.5)) of the float centre; Pillow
            # (12.2, Draw.c) truncates the box toward zero and fills it, except a box collapsed to a single point
            fx = centres(persons[p, :, 0], img_wh[0], W, center_mode, False)
            fy = centres(persons[p, :, 1], img_wh[1], H, center_mode, False)
            for k in range(J):
                if v[k] > 0:
                    x0, y0, x1, y1 = int(fx[k] - 0.5), int(fy[k] - 0.5), int(fx[k] + 0.5), int(fy[k] + 0.5)
                    if x1 > x0 or y1 > y0:
                        for yy in range(y0, y1 + 1):
                            for xx in range(x0, x1 + 1):
                                _put(canvas, xx, yy, k + 1)
        elif draw_points:
            for k in range(J):
                if v[k] > 0:
                    draw_point(canvas, x[k], y[k], k + 1)
        if draw_lines:
            for i, (a, b) in enumerate(limbs):
                if v[a] > 0 and v[b] > 0:
                    draw_line(canvas, x[a], y[a], x[b], y[b], line_value if line_value > 0 else (i if line_value < 0 else i + 1))
    return canvas.astype(np.int64)


def mpii_points_to_dense(points, J=16):
    """hourglass_compare.py:691-703 (= train.py:655-667): the sparse `annopoints.point` records of one MPII person ->
    points_rect[J, 3]; is_visible == 0 -> 0, anything else (1, or a missing / empty field) -> 1; later records of the same
    id overwrite earlier ones.  points: iterable of (id, x, y, is_visible)."""
    out = np.zeros([J, 3])
    for pid, x, y, vis in points:
        out[int(pid)] = [x, y, 0 if vis == 0 else 1]
    return out


def coco_persons_to_dense(persons, J=17):
    """try_with_torch.py:103-113: every `label['keypoints']` list of one image as [P, J, 3] (x, y, v)."""
    return np.asarray(persons, dtype=np.float64).reshape(-1, J, 3)
