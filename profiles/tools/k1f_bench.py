"""Times the features path on device-resident network-resolution maps (fused K1f + K2 + K3)
for the library variants under build/k1f_variants/ (or the default library)."""
import ctypes
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT))
from offsetguided_b200 import _lib, engine      # noqa: E402
from offsetguided_b200 import config as cfg     # noqa: E402
import bench                                    # noqa: E402


def run(path, hmp, omp, flip):
    lib = ctypes.CDLL(path)
    for nm, (restype, argtypes) in _lib.SIGNATURES.items():
        fn = getattr(lib, nm)
        fn.restype, fn.argtypes = restype, argtypes
    _lib._lib = lib
    skel = cfg.COCO_PERSON_SKELETON
    eng = engine.DecoderEngine(17, skel, topk=32, thre_hmp=0.04, dist_max=40, use_scale=True, person_thre=0.04)
    eng.enable_stage_timing(True)
    kp = cfg.heatmap_hflip(cfg.COCO_KEYPOINTS)
    fl, rs = cfg.offset_hflip(cfg.COCO_KEYPOINTS, skel)
    tables = (kp, fl, rs) if flip else None
    ts = []
    for i in range(13):
        poses = eng.decode_features(hmp, omp, 4, 4, 'bicubic', tables)
        if i >= 3:
            ts.append(eng.last_stage_times_ms())
    out = {'variant': os.path.basename(path), 'flip': flip, 'persons': sum(len(p) for p in poses)}
    for k in ('k1_stream', 'k1_select', 'k2', 'k3', 'd2h'):
        out[k + '_us'] = round(1e3 * float(np.mean([t[k] for t in ts])), 1)
    eng.close()
    return out


def main():
    paths = sorted(glob.glob(os.path.join(ROOT, 'build', 'k1f_variants', '*.so'))) or [_lib.LIB_PATH]
    for flip in (True, False):
        hmp, omp = bench.lowres_inputs(5000, 64, 640, flip)
        hmp, omp = torch.from_numpy(hmp).cuda(), torch.from_numpy(omp).cuda()
        for p in paths:
            print(json.dumps(run(p, hmp, omp, flip)))


if __name__ == '__main__':
    main()
