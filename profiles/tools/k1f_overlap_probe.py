"""n = 64 product path, 8 calls in flight, for a few block-kernel grid caps (run once per setting:
the knobs are read at first launch):  OG_K1F_CTAS_PER_SM=2 OG_K1F_WAVES=1 python profiles/tools/k1f_overlap_probe.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch            # noqa: E402
import bench            # noqa: E402
from offsetguided_b200 import config as cfg          # noqa: E402
from offsetguided_b200.engine import DecoderEngine   # noqa: E402

skel = cfg.COCO_PERSON_SKELETON
tables = (cfg.heatmap_hflip(cfg.COCO_KEYPOINTS),) + tuple(cfg.offset_hflip(cfg.COCO_KEYPOINTS, skel))
edge = int(sys.argv[1]) if len(sys.argv) > 1 else 640
hmp, omp = bench.lowres_inputs(5000, 64, edge, True)
bufs = [(torch.from_numpy(hmp).cuda(), torch.from_numpy(omp).cuda()) for _ in range(2)]
eng = DecoderEngine(17, skel, topk=32, thre_hmp=0.04, dist_max=40, use_scale=True, person_thre=0.04)
plans = [eng.plan_features(h, o, 4, 4, 'bicubic', tables) for h, o in bufs]
for depth in (2, 8):
    for i in range(3 * 8):
        plans[i % 2].launch()
        if i >= depth - 1:
            eng.fetch()
    while eng.pending:
        eng.fetch()
    torch.cuda.synchronize()
    steps = 2000
    t0 = time.perf_counter()
    for i in range(steps):
        plans[i % 2].launch()
        if i >= depth - 1:
            eng.fetch()
    while eng.pending:
        eng.fetch()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({'edge': edge, 'depth': depth, 'ctas_per_sm': os.environ.get('OG_K1F_CTAS_PER_SM', 'all'),
                      'waves': os.environ.get('OG_K1F_WAVES', '4'), 'ms_per_step': round(1e3 * dt / steps, 5),
                      'images_per_s': round(64 * steps / dt)}))
