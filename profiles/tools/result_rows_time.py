"""Result rows (back-projection + COCO formatting) for 64 and 256 images: the GPU epilogue
(PostProcess.generate_results -> arrays) against the host-numpy results.coco_results and the
reference's loops (oracle restatement), per batch."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402
import torch            # noqa: E402
import bench            # noqa: E402
from offsetguided_b200 import results                 # noqa: E402
from oracle import ref_oracle as ro                   # noqa: E402


class A(object):
    workload, batch, long_edge, no_flip = 'cfg2', 0, 0, False


def main():
    w = bench.workload(A())
    for n in (64, 256):
        post = bench.make_post(w, n)
        hmp, omp = bench.lowres_inputs(5000, n, 640, True, w)
        feats = [[[torch.from_numpy(hmp).cuda()], [[]], [[]]], [[torch.from_numpy(omp).cuda()], [[]], [[]]]]
        rng = np.random.RandomState(1)
        metas = [{'offset': np.array([-float(rng.randint(0, 90)), -float(rng.randint(0, 60))]),
                  'scale': np.array([rng.uniform(0.5, 1.5)] * 2), 'hflip': False, 'image_id': i} for i in range(n)]
        for _ in range(5):
            poses = post.generate_poses(feats, flip_test=True)
            post.generate_results(feats, metas, flip_test=True)
        reps = 50
        t0 = time.perf_counter()
        for _ in range(reps):
            poses = post.generate_poses(feats, flip_test=True)
        t_plain = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            poses, kp, sc, img = post.generate_results(feats, metas, flip_test=True)
        t_gpu = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(5):
            results.coco_results(poses, metas)
        t_numpy = (time.perf_counter() - t0) / 5
        t0 = time.perf_counter()
        for _ in range(3):
            ro.coco_result_rows(poses, metas)
        t_loops = (time.perf_counter() - t0) / 3
        t0 = time.perf_counter()
        for _ in range(5):
            results.rows_from_arrays(kp, sc, img, metas)
        t_dicts = (time.perf_counter() - t0) / 5
        print(json.dumps({'images': n, 'persons': int(sum(len(p) for p in poses)),
                          'generate_poses_ms': round(1e3 * t_plain, 4),
                          'generate_results_ms (poses + result arrays, GPU epilogue)': round(1e3 * t_gpu, 4),
                          'added_by_result_arrays_ms': round(1e3 * (t_gpu - t_plain), 4),
                          'results.coco_results_ms (host numpy + dict rows)': round(1e3 * t_numpy, 3),
                          'reference_loops_ms (oracle restatement)': round(1e3 * t_loops, 3),
                          'rows_from_arrays_ms (dict rows from the GPU arrays)': round(1e3 * t_dicts, 3)}))


if __name__ == '__main__':
    main()
