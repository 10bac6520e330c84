"""Host-side timing of the pipelined device-resident features path: how long the launch call
and the fetch take per step, with and without per-stage event timing."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np      # noqa: E402
import torch            # noqa: E402
import bench            # noqa: E402
from offsetguided_b200 import config as cfg          # noqa: E402
from offsetguided_b200.engine import DecoderEngine   # noqa: E402


def main():
    skel = cfg.COCO_PERSON_SKELETON
    tables = (cfg.heatmap_hflip(cfg.COCO_KEYPOINTS),) + tuple(cfg.offset_hflip(cfg.COCO_KEYPOINTS, skel))
    hmp, omp = bench.lowres_inputs(5000, 64, 640, True)
    th, to = torch.from_numpy(hmp).cuda(), torch.from_numpy(omp).cuda()
    for timing in (True, False):
        eng = DecoderEngine(17, skel, topk=32, thre_hmp=0.04, dist_max=40, use_scale=True, person_thre=0.04)
        eng.enable_stage_timing(timing)
        for _ in range(5):
            eng.decode_features(th, to, 4, 4, 'bicubic', tables)
        steps = 200
        t_launch, t_fetch = [], []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        depth = int(os.environ.get('PIPE_DEPTH', '2'))           # decode calls in flight
        for _ in range(depth - 1):
            eng.decode_features(th, to, 4, 4, 'bicubic', tables, fetch=False)
        for _ in range(steps - (depth - 1)):
            a = time.perf_counter()
            eng.decode_features(th, to, 4, 4, 'bicubic', tables, fetch=False)
            b = time.perf_counter()
            eng.fetch(64)
            c = time.perf_counter()
            t_launch.append(b - a)
            t_fetch.append(c - b)
        for _ in range(depth - 1):
            eng.fetch(64)
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
        print(json.dumps({'stage_timing': timing, 'depth': depth, 'ms_per_step': 1e3 * total / steps,
                          'launch_call_us': 1e6 * float(np.median(t_launch)),
                          'fetch_call_us': 1e6 * float(np.median(t_fetch)),
                          'images_per_s': 64 * steps / total}))
        eng.close()


if __name__ == '__main__':
    main()
