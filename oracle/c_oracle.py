"""TEST INFRASTRUCTURE — ctypes front end of oracle/og_oracle.c (the fast CPU oracle).
Same call signatures as oracle/ref_oracle.py.  Never imported by the product package."""
import ctypes
import os

import numpy as np

from . import build_oracle

F32 = np.float32
_lib = None


def load():
    global _lib
    if _lib is None:
        path = build_oracle.build()
        _lib = ctypes.CDLL(path)
        _lib.oc_num_threads.restype = ctypes.c_int
        _lib.oc_group_image.restype = ctypes.c_int
        # torchrun exports OMP_NUM_THREADS=1; the CPU baseline uses every host core
        _lib.oc_set_num_threads(ctypes.c_int(os.cpu_count() or 1))
    return _lib


def num_threads():
    return int(load().oc_num_threads())


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i32(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.int32))


def flip_augment(hmps, offs, kp_flips, limb_flips, limb_reserve):
    lib = load()
    hmps = np.ascontiguousarray(hmps, dtype=F32)
    offs = np.ascontiguousarray(offs, dtype=F32)
    n2, c, h, w = hmps.shape
    l = offs.shape[1] // 2
    n = n2 // 2
    out_h = np.empty((n, c, h, w), F32)
    out_o = np.empty((n, 2 * l, h, w), F32)
    kp, lf, lr = _i32(kp_flips), _i32(limb_flips), _i32(limb_reserve)
    lib.oc_flip_fuse(_p(hmps), _p(offs), _p(kp), _p(lf), _p(lr), ctypes.c_int(len(lr)), n, c, l, h, w,
                     _p(out_h), _p(out_o))
    return out_h, out_o


def resize(x, scale, mode):
    lib = load()
    x = np.ascontiguousarray(x, dtype=F32)
    h, w = x.shape[-2:]
    planes = int(np.prod(x.shape[:-2]))
    out = np.empty(x.shape[:-2] + (h * scale, w * scale), F32)
    lib.oc_resize(_p(x), _p(out), planes, h, w, int(scale), 1 if mode == 'bicubic' else 0)
    return out


def joint_dets(hmps, k):
    lib = load()
    hmps = np.ascontiguousarray(hmps, dtype=F32)
    n, c, h, w = hmps.shape
    s = np.empty((n, c, k), F32)
    i = np.empty((n, c, k), np.int64)
    lib.oc_nms_topk(_p(hmps), n * c, h, w, int(k), _p(s), _p(i))
    return s, i, i // w, i % w


def generate_limbs(hmps_hr, offs_hr, skeleton, topk, thre_hmp, min_len, hmp_s=4, off_s=4,
                   scmps_hr=None, return_dets=False):
    lib = load()
    hmps_hr = np.ascontiguousarray(hmps_hr, dtype=F32)
    offs_hr = np.ascontiguousarray(offs_hr, dtype=F32)
    assert hmps_hr.shape[-2:] == offs_hr.shape[-2:], 'spatial resolution should be equal'
    n, c, h, w = hmps_hr.shape
    l = len(skeleton)
    dets = joint_dets(hmps_hr, topk)
    fr, to = _i32([a for a, _ in skeleton]), _i32([b for _, b in skeleton])
    limbs = np.empty((n, l, topk, 13), F32)
    sc = np.ascontiguousarray(scmps_hr, dtype=F32) if scmps_hr is not None else None
    lib.oc_limbs(_p(dets[0]), _p(dets[1]), _p(offs_hr), _p(sc) if sc is not None else None,
                 n, c, l, int(topk), h, w, _p(fr), _p(to), ctypes.c_float(thre_hmp),
                 ctypes.c_float(min_len), ctypes.c_float(off_s / hmp_s), _p(limbs))
    if return_dets:
        return limbs, dets
    return limbs


def group_batch(limbs, skeleton, n_keypoints, person_thre, sort_dim=2, dist_max=10, use_scale=False):
    lib = load()
    limbs = np.ascontiguousarray(limbs, dtype=F32)
    n, l, k, _ = limbs.shape
    assert l == len(skeleton), 'check the skeleton config and input limbs Tensor'
    fr, to = _i32([a for a, _ in skeleton]), _i32([b for _, b in skeleton])
    out = np.empty((n, l * k, n_keypoints, 6), F32)
    counts = np.zeros((n,), np.int32)
    lib.oc_group_batch(_p(limbs), n, int(n_keypoints), l, k, _p(fr), _p(to), ctypes.c_float(dist_max),
                       1 if use_scale else 0, ctypes.c_double(person_thre), int(sort_dim), _p(out),
                       _p(counts))
    return [out[i, :counts[i]].copy() for i in range(n)]


def group_skeletons(limbs, skeleton, n_keypoints, person_thre, sort_dim=2, dist_max=10,
                    use_scale=False):
    return group_batch(np.asarray(limbs)[None], skeleton, n_keypoints, person_thre, sort_dim,
                       dist_max, use_scale)[0]


def generate_poses(hmps, offs, skeleton, n_keypoints, *, topk, thre_hmp, min_len, person_thre,
                   sort_dim=2, dist_max=20, use_scale=True, hmp_stride=4, off_stride=4,
                   resize_mode='bicubic', flip_test=False, kp_flips=None, limb_flips=None,
                   limb_reserve=None, return_limbs=False, stage_seconds=None):
    """``stage_seconds`` (a dict) accumulates the wall time of every stage, for the CPU baseline
    of bench.py: flip fusion / resize / NMS + top-K + limbs / grouping."""
    import time
    t = [time.perf_counter()]

    def lap(name):
        t.append(time.perf_counter())
        if stage_seconds is not None:
            stage_seconds[name] = stage_seconds.get(name, 0.0) + t[-1] - t[-2]
    if flip_test:
        hmps, offs = flip_augment(hmps, offs, kp_flips, limb_flips, limb_reserve)
    lap('flip_fusion')
    hmps_hr = resize(hmps, hmp_stride, resize_mode) if hmp_stride > 1 else hmps
    offs_hr = resize(offs, off_stride, 'bilinear') if off_stride > 1 else offs
    lap('resize')
    limbs = generate_limbs(hmps_hr, offs_hr, skeleton, topk, thre_hmp, min_len, hmp_stride, off_stride)
    lap('nms_topk_limbs')
    poses = group_batch(limbs, skeleton, n_keypoints, person_thre, sort_dim, dist_max, use_scale)
    lap('grouping')
    if return_limbs:
        return poses, limbs
    return poses
