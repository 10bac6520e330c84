"""K3 tuning sweep (GPU box): variants of the library built with different OG_K3_THREADS,
timed with the library's own stage events on two workloads."""
import ctypes
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from offsetguided_b200 import _lib, engine      # noqa: E402
from offsetguided_b200 import config as cfg     # noqa: E402
from oracle import scenes                       # noqa: E402


def run(path, name, heat, offs, skel, c, k, reps=10):
    lib = ctypes.CDLL(path)
    for nm, (restype, argtypes) in _lib.SIGNATURES.items():
        fn = getattr(lib, nm)
        fn.restype, fn.argtypes = restype, argtypes
    _lib._lib = lib                      # engine uses this library
    eng = engine.DecoderEngine(c, skel, topk=k, thre_hmp=0.04, dist_max=40, use_scale=True, person_thre=0.04)
    eng.enable_stage_timing(True)
    ts, persons = [], 0
    for i in range(reps + 3):
        poses = eng.decode_maps(heat, offs)
        if i >= 3:
            ts.append(eng.last_stage_times_ms())
        persons = sum(len(p) for p in poses)
    out = {'variant': os.path.basename(path), 'workload': name, 'persons': persons,
           'k3_us': round(1e3 * float(np.mean([t['k3'] for t in ts])), 2),
           'k2_us': round(1e3 * float(np.mean([t['k2'] for t in ts])), 2),
           'chk': float(sum(float(p.sum()) for p in poses))}
    eng.close()
    return out


def main():
    skel = cfg.COCO_PERSON_SKELETON
    h1, o1 = scenes.synth_hires_batch(1000, 8, 6, 640, 640, skel)
    h1 = torch.from_numpy(h1).cuda().repeat(8, 1, 1, 1).contiguous()
    o1 = torch.from_numpy(o1).cuda().repeat(8, 1, 1, 1).contiguous()
    h2, o2 = scenes.synth_hires_batch(2000, 8, 20, 640, 640, skel)
    h2 = torch.from_numpy(h2).cuda().repeat(4, 1, 1, 1).contiguous()
    o2 = torch.from_numpy(o2).cuda().repeat(4, 1, 1, 1).contiguous()
    for path in sorted(glob.glob(os.path.join(ROOT, 'build', 'k3_sweep', '*.so'))):
        print(json.dumps(run(path, '64 img x 6 persons K=32', h1, o1, skel, 17, 32)))
        print(json.dumps(run(path, '32 img x 20 persons K=64', h2, o2, skel, 17, 64)))


if __name__ == '__main__':
    main()
