"""A few kernel-by-kernel decodes of the device-resident features path (for ncu launch lists):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv python profiles/tools/chain_once.py [n] [edge]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch            # noqa: E402
import bench            # noqa: E402
from offsetguided_b200 import config as cfg          # noqa: E402
from offsetguided_b200.engine import DecoderEngine   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
edge = int(sys.argv[2]) if len(sys.argv) > 2 else 640
skel = cfg.COCO_PERSON_SKELETON
tables = (cfg.heatmap_hflip(cfg.COCO_KEYPOINTS),) + tuple(cfg.offset_hflip(cfg.COCO_KEYPOINTS, skel))
hmp, omp = bench.lowres_inputs(5000, n, edge, True)
th, to = torch.from_numpy(hmp).cuda(), torch.from_numpy(omp).cuda()
eng = DecoderEngine(17, skel, topk=32, thre_hmp=0.04, dist_max=40, use_scale=True, person_thre=0.04)
eng.set_graph(False)
for _ in range(4):
    poses = eng.decode_features(th, to, 4, 4, 'bicubic', tables)
print(n, edge, sum(len(p) for p in poses), 'persons')
