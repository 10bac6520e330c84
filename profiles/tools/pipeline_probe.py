"""Device-resident features path (what evaluate.py calls after model(images)): ms per step and
host microseconds per launch / fetch call, for batch shards of 8 .. 64 images, 1 .. 8 calls in
flight, CUDA-graph replay on / off, through decode_features and through a prepared plan.

    python profiles/tools/pipeline_probe.py [long_edge] [n,n,...] [depth,depth,...]
"""
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


def run(eng, launch, fetch, n, depth, steps):
    for _ in range(10):
        launch()
        fetch()
    torch.cuda.synchronize()
    t_launch, t_fetch = [], []
    t0 = time.perf_counter()
    for _ in range(depth - 1):
        launch()
    for _ in range(steps - (depth - 1)):
        a = time.perf_counter()
        launch()
        b = time.perf_counter()
        fetch()
        c = time.perf_counter()
        t_launch.append(b - a)
        t_fetch.append(c - b)
    for _ in range(depth - 1):
        fetch()
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    return {'ms_per_step': round(1e3 * total / steps, 5), 'images_per_s': round(n * steps / total),
            'launch_us': round(1e6 * float(np.median(t_launch)), 2),
            'fetch_us': round(1e6 * float(np.median(t_fetch)), 2)}


def main():
    edge = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    skel = cfg.COCO_PERSON_SKELETON
    tables = (cfg.heatmap_hflip(cfg.COCO_KEYPOINTS),) + tuple(cfg.offset_hflip(cfg.COCO_KEYPOINTS, skel))
    hmp, omp = bench.lowres_inputs(5000, 64, edge, True)
    eng = DecoderEngine(17, skel, topk=32, thre_hmp=0.04, dist_max=40, use_scale=True, person_thre=0.04)
    ns = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [8, 16, 32, 64]
    depths = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [1, 2, 4, 8, 16]
    for n in ns:
        sel = list(range(n)) + list(range(64, 64 + n))
        th, to = torch.from_numpy(hmp[sel]).cuda(), torch.from_numpy(omp[sel]).cuda()
        # stage times, kernel by kernel
        eng.enable_stage_timing(True)
        for _ in range(5):
            eng.decode_features(th, to, 4, 4, 'bicubic', tables)
        print(json.dumps({'n': n, 'edge': edge, 'stage_ms': {k: round(v, 5) for k, v in eng.last_stage_times_ms().items()}}))
        eng.enable_stage_timing(False)
        for graph in (True, False):
            eng.set_graph(graph)
            for depth in depths:
                steps = 400
                r = run(eng, lambda: eng.decode_features(th, to, 4, 4, 'bicubic', tables, fetch=False),
                        lambda: eng.fetch(), n, depth, steps)
                r.update({'n': n, 'edge': edge, 'graph': graph, 'depth': depth, 'api': 'decode_features'})
                print(json.dumps(r))
                if graph:
                    plan = eng.plan_features(th, to, 4, 4, 'bicubic', tables)
                    r = run(eng, plan.launch, plan.fetch, n, depth, steps)
                    r.update({'n': n, 'edge': edge, 'graph': graph, 'depth': depth, 'api': 'plan'})
                    print(json.dumps(r))
        eng.set_graph(True)
    print(json.dumps({'graph_counts': eng.graph_counts}))


if __name__ == '__main__':
    main()
