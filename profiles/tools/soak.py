"""Soak test: thousands of pipelined decode calls on all three entry paths (full-resolution maps,
device-resident network-resolution maps, pinned host maps), every result compared with the first
one.  Catches rare ordering bugs between the caller's stream, the handle's stream and the copy
stream that single-shot tests cannot."""
import json
import sys
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np      # noqa: E402
import torch            # noqa: E402
import bench            # noqa: E402
from offsetguided_b200 import config as cfg          # noqa: E402
from offsetguided_b200.engine import DecoderEngine   # noqa: E402
from oracle import scenes                            # noqa: E402


def same(a, b):
    return len(a) == len(b) and all(x.shape == y.shape and np.array_equal(x, y) for x, y in zip(a, b))


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    skel = cfg.COCO_PERSON_SKELETON
    tables = (cfg.heatmap_hflip(cfg.COCO_KEYPOINTS),) + tuple(cfg.offset_hflip(cfg.COCO_KEYPOINTS, skel))
    n = 16
    hmp, omp = bench.lowres_inputs(4242, n, 640, True)
    heat, offs = scenes.synth_hires_batch(99, n, 6, 640, 640, skel)
    th_d, to_d = torch.from_numpy(hmp).cuda(), torch.from_numpy(omp).cuda()
    th_h, to_h = torch.from_numpy(hmp).pin_memory(), torch.from_numpy(omp).pin_memory()
    heat_d, offs_d = torch.from_numpy(heat).cuda(), torch.from_numpy(offs).cuda()
    eng = DecoderEngine(17, skel, topk=32, thre_hmp=0.04, dist_max=40, use_scale=True, person_thre=0.04)
    calls = {
        'maps': lambda fetch: eng.decode_maps(heat_d, offs_d, fetch=fetch),
        'features_dev': lambda fetch: eng.decode_features(th_d, to_d, 4, 4, 'bicubic', tables, fetch=fetch),
        'features_host': lambda fetch: eng.decode_features(th_h, to_h, 4, 4, 'bicubic', tables, fetch=fetch),
    }
    ref = {k: f(True) for k, f in calls.items()}
    assert same(ref['features_dev'], ref['features_host'])
    bad = {k: 0 for k in calls}
    order = list(calls)
    rng = np.random.RandomState(0)
    t0 = time.time()
    pending = []
    for it in range(iters):
        k = order[rng.randint(len(order))]
        calls[k](False)
        pending.append(k)
        if len(pending) == 3:                       # three calls in flight, mixed paths
            kk = pending.pop(0)
            if not same(eng.fetch(n), ref[kk]):
                bad[kk] += 1
    while pending:
        kk = pending.pop(0)
        if not same(eng.fetch(n), ref[kk]):
            bad[kk] += 1
    print(json.dumps({'iterations': iters, 'mismatches': bad, 'seconds': round(time.time() - t0, 1),
                      'persons': {k: sum(len(p) for p in v) for k, v in ref.items()},
                      'fused_redos': eng.fused_redo_count}))
    assert not any(bad.values())


if __name__ == '__main__':
    main()
