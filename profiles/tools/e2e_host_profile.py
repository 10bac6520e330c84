"""Where the end-to-end step of the host API goes: host time before the call returns, wait for the
result, host-side unpacking; for OG_HOST_CHUNKS = 1, 2, 4, 8 (copy / decode pipeline depth)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    import bench
    from offsetguided_b200 import config as cfg
    from offsetguided_b200.engine import DecoderEngine
    skel = cfg.COCO_PERSON_SKELETON
    kp = cfg.heatmap_hflip(cfg.COCO_KEYPOINTS)
    fl, rs = cfg.offset_hflip(cfg.COCO_KEYPOINTS, skel)
    hmp, omp = bench.lowres_inputs(5000, 64, 640, True)
    th, to = torch.from_numpy(hmp).pin_memory(), torch.from_numpy(omp).pin_memory()
    eng = DecoderEngine(17, skel, topk=32, thre_hmp=0.04, dist_max=40, use_scale=True, person_thre=0.04)
    eng.enable_stage_timing(True)
    for _ in range(5):
        eng.decode_features(th, to, 4, 4, 'bicubic', (kp, fl, rs))
    t_call, t_wait, t_total, stages = [], [], [], []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = eng.decode_features(th, to, 4, 4, 'bicubic', (kp, fl, rs), fetch=False)
        t1 = time.perf_counter()
        poses = eng.fetch(n)
        t2 = time.perf_counter()
        t_call.append(t1 - t0)
        t_wait.append(t2 - t1)
        t_total.append(t2 - t0)
        stages.append(eng.last_stage_times_ms())
    out = {'chunks': os.environ.get('OG_HOST_CHUNKS', 'default'),
           'call_ms': 1e3 * float(np.median(t_call)), 'fetch_ms': 1e3 * float(np.median(t_wait)),
           'total_ms': 1e3 * float(np.median(t_total)),
           'stage_ms': {k: round(float(np.median([s[k] for s in stages])), 4) for k in stages[0]},
           'persons': sum(len(p) for p in poses)}
    print(json.dumps(out))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        child()
    else:
        for chunks in ('1', '2', '4', '8'):
            env = dict(os.environ, OG_HOST_CHUNKS=chunks)
            subprocess.run([sys.executable, os.path.abspath(__file__), 'child'], env=env, check=False)
