"""TEST INFRASTRUCTURE — synthetic multi-person scenes for the decoder parity tests.

Only tests/, __graft_entry__.smoke() and bench.py may import this module; the
product package (offsetguided_b200/) never does.

The reference renders its network targets with ``encoder/heatmap.py`` and
``encoder/offset.py``.  Those files cannot travel to the GPU box, so this module
restates the two renderers (numpy) for use at run time there;
``tests/golden/make_golden.py`` checks the restatement bit-for-bit against the
reference encoder inside the build container and commits reference-rendered
fixtures.

  * persons: canonical template pose of reference config/coco_data.py:184-202
    placed at random centres / scales (SURVEY.md 8d "synthetic scene generator");
  * ``render_heatmaps``  follows encoder/heatmap.py:100-178 (Gaussian, sigma in
    input pixels, clip threshold, max-overlap, window of ``gaussian_size`` cells);
  * ``render_offsets``   follows encoder/offset.py:76-197 (7x7 fill around the
    from-joint, shortest-vector-wins overlap, +inf background).
"""
import math

import numpy as np

# (x, y) of the 17 COCO joints in a y-up unit frame
TEMPLATE_COCO = np.array([
    [0.0, 9.3], [-0.5, 9.7], [0.5, 9.7], [-1.0, 9.5], [1.0, 9.5],
    [-2.0, 8.0], [2.0, 8.0], [-2.5, 6.0], [2.5, 6.2], [-2.5, 4.0],
    [2.5, 4.2], [-1.8, 4.0], [1.8, 4.0], [-2.0, 2.0], [2.0, 2.1],
    [-2.0, 0.0], [2.0, 0.1]], dtype=np.float64)

# 14 CrowdPose joints (shoulders, elbows, wrists, hips, knees, ankles, head, neck)
TEMPLATE_CROWDPOSE = np.array([
    [-2.0, 8.0], [2.0, 8.0], [-2.5, 6.0], [2.5, 6.2], [-2.5, 4.0], [2.5, 4.2],
    [-1.8, 4.0], [1.8, 4.0], [-2.0, 2.0], [2.0, 2.1], [-2.0, 0.0], [2.0, 0.1],
    [0.0, 10.2], [0.0, 8.6]], dtype=np.float64)


def make_persons(rng, n_persons, width, height, template=TEMPLATE_COCO,
                 scale_range=(10.0, 24.0), jitter=0.15, drop_prob=0.0):
    """(P, C, 4) float32 array of (x, y, v, joint_scale) in input pixels.

    ``rng`` is a ``numpy.random.RandomState``.  Every joint lies strictly inside
    the image with a margin, so x > 0 and y > 0 (reference group.py:74-75 drops
    peaks in row/column 0)."""
    c = template.shape[0]
    out = np.zeros((n_persons, c, 4), dtype=np.float32)
    tx, ty = template[:, 0], template[:, 1]
    for p in range(n_persons):
        s = rng.uniform(*scale_range)
        w_half = 2.5 * s + 12
        top = 10.2 * s + 12
        cx = rng.uniform(w_half, width - w_half)
        feet_y = rng.uniform(top, height - 12)
        jit = rng.normal(0.0, jitter, size=(c, 2)) * s
        x = cx + s * tx + jit[:, 0]
        y = feet_y - s * ty + jit[:, 1]
        out[p, :, 0] = np.clip(x, 6, width - 7)
        out[p, :, 1] = np.clip(y, 6, height - 7)
        out[p, :, 2] = 2.0
        out[p, :, 3] = s / 2.0
        if drop_prob > 0:
            dropped = rng.uniform(size=c) < drop_prob
            out[p, dropped, 2] = 0.0
    return out


def render_heatmaps(persons, width, height, stride=4, sigma=7.0, clip_thre=0.01):
    """(C, H/stride, W/stride) float32 Gaussian heatmaps."""
    c = persons.shape[1]
    ow, oh = width // stride, height // stride
    dsig = 2 * sigma * sigma
    gsize = 2 * math.ceil(math.sqrt(-dsig * math.log(clip_thre)) / stride)
    gx = np.arange(ow) * stride + stride / 2 - 0.5
    gy = np.arange(oh) * stride + stride / 2 - 0.5
    dsig32 = np.array([dsig]).astype(np.float32)
    hm = np.zeros((c, oh, ow), dtype=np.float32)
    for ch in range(c):
        for j in persons[persons[:, ch, 2] > 0, ch]:
            x0 = int(round(j[0] / stride - gsize / 2))
            x1 = int(round(j[0] / stride + gsize / 2))
            y0 = int(round(j[1] / stride - gsize / 2))
            y1 = int(round(j[1] / stride + gsize / 2))
            if y1 < 0 or x1 < 0:
                continue
            x0, y0 = max(x0, 0), max(y0, 0)
            dx = gx[x0:x1].astype(np.float32) - j[0]
            dy = gy[y0:y1].astype(np.float32) - j[1]
            ex = np.exp(-dx ** 2 / dsig32)
            ey = np.exp(-dy ** 2 / dsig32)
            g = np.outer(ey, ex)
            g[g < clip_thre] = 0
            patch = hm[ch, y0:y1, x0:x1]
            np.maximum(patch, g, out=patch)
    return hm


def render_offsets(persons, width, height, skeleton, stride=4, fill=7):
    """(2L, H/stride, W/stride) float32 guiding-offset maps, +inf background."""
    ow, oh = width // stride, height // stride
    gx = np.arange(ow) * stride + stride / 2 - 0.5
    gy = np.arange(oh) * stride + stride / 2 - 0.5
    om = np.full((2 * len(skeleton), oh, ow), np.inf, dtype=np.float32)
    for l, (fr, to) in enumerate(skeleton):
        vis = (persons[:, fr, 2] > 0) & (persons[:, to, 2] > 0)
        for j1, j2 in zip(persons[vis, fr], persons[vis, to]):
            x0 = int(round(j1[0] / stride - fill / 2))
            x1 = int(round(j1[0] / stride + fill / 2))
            y0 = int(round(j1[1] / stride - fill / 2))
            y1 = int(round(j1[1] / stride + fill / 2))
            if y1 < 0 or x1 < 0:
                continue
            x0, y0 = max(x0, 0), max(y0, 0)
            ox = j2[0] - gx[x0:x1].astype(np.float32)
            oy = j2[1] - gy[y0:y1].astype(np.float32)
            mx = np.broadcast_to(ox[None, :], (oy.shape[0], ox.shape[0]))
            my = np.broadcast_to(oy[:, None], (oy.shape[0], ox.shape[0]))
            new_len = np.linalg.norm(np.stack((mx, my), axis=-1), axis=-1)
            px = om[2 * l, y0:y1, x0:x1]
            py = om[2 * l + 1, y0:y1, x0:x1]
            old_len = np.linalg.norm(np.stack((px, py), axis=-1), axis=-1)
            closer = new_len < old_len
            px[closer] = mx[closer]
            py[closer] = my[closer]
    return om


def mirror_persons(persons, width, kp_flips):
    """Persons as seen in the W-flipped image (x -> W-1-x, left/right swapped)."""
    out = persons[:, kp_flips, :].copy()
    out[:, :, 0] = (width - 1) - out[:, :, 0]
    return out


def render_scene(rng, n_persons, width, height, skeleton, template=TEMPLATE_COCO,
                 stride=4, noise=0.0, scale_range=(10.0, 24.0), drop_prob=0.0):
    """One synthetic network output: (hmp (C,h,w), omp (2L,h,w), persons)."""
    persons = make_persons(rng, n_persons, width, height, template,
                           scale_range=scale_range, drop_prob=drop_prob)
    hmp = render_heatmaps(persons, width, height, stride)
    omp = render_offsets(persons, width, height, skeleton, stride)
    omp[~np.isfinite(omp)] = 0.0
    if noise > 0:
        hmp = hmp + rng.uniform(0.0, noise, size=hmp.shape).astype(np.float32)
    return hmp, omp, persons


def render_batch(seed, n_images, n_persons, width, height, skeleton,
                 template=TEMPLATE_COCO, stride=4, noise=0.0,
                 scale_range=(10.0, 24.0), drop_prob=0.0):
    """Batch of scenes: hmp (N,C,h,w), omp (N,2L,h,w); image i uses seed+i."""
    hs, os_ = [], []
    for i in range(n_images):
        rng = np.random.RandomState(seed + i)
        h, o, _ = render_scene(rng, n_persons, width, height, skeleton, template,
                               stride, noise, scale_range, drop_prob)
        hs.append(h)
        os_.append(o)
    return np.stack(hs), np.stack(os_)


def synth_hires_batch(seed, n_images, n_persons, width, height, skeleton,
                      n_channels=17, sigma=5.0, noise=0.02):
    """Cheap full-resolution maps for throughput runs (no x4 resize involved):
    heat (N,C,H,W) with Gaussian peaks + uniform noise floor, offs (N,2L,H,W)
    holding the exact from->to vector in a 9x9 window round each from-joint."""
    template = TEMPLATE_COCO if n_channels == 17 else TEMPLATE_CROWDPOSE
    heat = np.zeros((n_images, n_channels, height, width), dtype=np.float32)
    offs = np.zeros((n_images, 2 * len(skeleton), height, width), dtype=np.float32)
    r = int(3 * sigma)
    ax = np.arange(-r, r + 1, dtype=np.float32)
    for i in range(n_images):
        rng = np.random.RandomState(seed + i)
        persons = make_persons(rng, n_persons, width, height, template,
                               scale_range=(width / 64.0, width / 27.0))
        if noise > 0:
            heat[i] = rng.uniform(0.0, noise, size=heat[i].shape).astype(np.float32)
        for p in persons:
            amp = rng.uniform(0.5, 1.0, size=n_channels)
            for c in range(n_channels):
                x, y = p[c, 0], p[c, 1]
                xi, yi = int(round(x)), int(round(y))
                gxv = np.exp(-((ax + xi - x) ** 2) / (2 * sigma * sigma))
                gyv = np.exp(-((ax + yi - y) ** 2) / (2 * sigma * sigma))
                g = (amp[c] * np.outer(gyv, gxv)).astype(np.float32)
                ys, ye = max(yi - r, 0), min(yi + r + 1, height)
                xs, xe = max(xi - r, 0), min(xi + r + 1, width)
                patch = heat[i, c, ys:ye, xs:xe]
                np.maximum(patch, g[ys - yi + r:ye - yi + r, xs - xi + r:xe - xi + r],
                           out=patch)
            for l, (fr, to) in enumerate(skeleton):
                x, y = p[fr, 0], p[fr, 1]
                xi, yi = int(round(x)), int(round(y))
                ys, ye = max(yi - 4, 0), min(yi + 5, height)
                xs, xe = max(xi - 4, 0), min(xi + 5, width)
                offs[i, 2 * l, ys:ye, xs:xe] = p[to, 0] - np.arange(xs, xe, dtype=np.float32)[None, :]
                offs[i, 2 * l + 1, ys:ye, xs:xe] = p[to, 1] - np.arange(ys, ye, dtype=np.float32)[:, None]
    return heat, offs
