"""K1 tuning sweep (run on the GPU box): builds variants of the library with different
OG_K1_* knobs (prebuilt here, shipped in build/k1_variants/) and times og_nms_topk_f32 on the
bench workload with CUDA events.  Output: one line per variant."""
import ctypes
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from offsetguided_b200 import _lib          # noqa: E402
from oracle import scenes                   # noqa: E402
from offsetguided_b200 import config as cfg  # noqa: E402


def time_variant(path, heat, reps=20):
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = restype, argtypes
    n, c, h, w = heat.shape
    fr = _lib.int32_array([a for a, _ in cfg.COCO_PERSON_SKELETON])
    to = _lib.int32_array([b for _, b in cfg.COCO_PERSON_SKELETON])
    conf = _lib.OgConfig(n_keypoints=17, n_limbs=19, limb_from=ctypes.cast(fr, _lib.c_int32_p),
                         limb_to=ctypes.cast(to, _lib.c_int32_p), topk=32, thre_hmp=0.04, min_len=0.5,
                         resize_factor=1.0, dist_max=40.0, use_scale=1, person_thre=0.04, sort_dim=2,
                         device=0, max_images=0)
    hd = ctypes.c_void_p()
    assert lib.og_create(ctypes.byref(conf), ctypes.byref(hd)) == 0, lib.og_last_error()
    s = torch.empty((n, c, 32), device='cuda')
    i = torch.empty((n, c, 32), dtype=torch.int32, device='cuda')
    cnt = torch.empty((n, c), dtype=torch.int32, device='cuda')
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run():
        st = lib.og_nms_topk_f32(hd, ctypes.c_void_p(heat.data_ptr()), n, h, w, ctypes.c_float(0.04),
                                 ctypes.c_void_p(s.data_ptr()), ctypes.c_void_p(i.data_ptr()),
                                 ctypes.c_void_p(cnt.data_ptr()), stream)
        assert st == 0, lib.og_last_error()
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    chk = (int(i.sum().item()), float(s.sum().item()), int(cnt.sum().item()))
    lib.og_destroy(hd)
    return ms, chk


def main():
    heat_np, _ = scenes.synth_hires_batch(1000, 8, 6, 640, 640, cfg.COCO_PERSON_SKELETON)
    heat = torch.from_numpy(heat_np).cuda().repeat(8, 1, 1, 1).contiguous()
    nbytes = heat.numel() * 4
    out = []
    for path in sorted(glob.glob(os.path.join(ROOT, 'build', 'k1_variants', '*.so'))):
        ms, chk = time_variant(path, heat)
        rec = {'variant': os.path.basename(path), 'ms': round(ms, 4), 'GBps': round(nbytes / ms / 1e6, 1),
               'check': chk}
        print(json.dumps(rec))
        out.append(rec)
    assert len({tuple(r['check']) for r in out}) <= 1, 'variants disagree'


if __name__ == '__main__':
    main()
