"""Soak test: thousands of pipelined decode calls on all entry paths (full-resolution maps,
device-resident network-resolution maps — direct call, prepared plan, a batch with a noisy plane
that is redone at fetch time — and pinned host maps), a random number of calls (up to 12) in
flight, every result compared with the first one.  Catches rare ordering bugs between the
caller's stream, the result slots' streams, graph replays and the copy stream that single-shot
tests cannot."""
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


def same_rows(a, b):
    """Result rows of two calls (engine.last_result_rows: rows in the order the row allocator handed
    them out, which differs from run to run) compared in image / rank order."""
    if a is None or b is None:
        return a is None and b is None
    from offsetguided_b200 import results
    return all(np.array_equal(x, y) for x, y in zip(results.result_arrays(a, 17), results.result_arrays(b, 17)))


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
    noisy = hmp.copy()
    noisy[3, 7] = np.random.RandomState(1).uniform(0, 1, size=hmp.shape[2:]).astype(np.float32)
    tn_d = torch.from_numpy(noisy).cuda()
    plan = eng.plan_features(th_d.clone(), to_d.clone(), 4, 4, 'bicubic', tables)
    rs = np.random.RandomState(2)
    sc_d = torch.from_numpy(rs.uniform(2, 9, size=hmp.shape).astype(np.float32)).cuda()          # keypoint-scale maps
    jo_d = torch.from_numpy(rs.uniform(-1, 1, size=(hmp.shape[0], 2) + hmp.shape[2:]).astype(np.float32)).cuda()
    frames = np.stack([np.array([3.0 * i, 2.0 * i, 1.25, 1.25]) for i in range(n)])            # offset x, y, scale x, y

    def plan_call(fetch):
        plan.launch()
        return eng.fetch() if fetch else None
    calls = {
        'maps': lambda fetch: eng.decode_maps(heat_d, offs_d, fetch=fetch),
        'features_dev': lambda fetch: eng.decode_features(th_d, to_d, 4, 4, 'bicubic', tables, fetch=fetch),
        'features_host': lambda fetch: eng.decode_features(th_h, to_h, 4, 4, 'bicubic', tables, fetch=fetch),
        'features_plan': plan_call,
        'features_heads': lambda fetch: eng.decode_features_heads(th_d, to_d, sc_d, jo_d, 4, 4, 'bicubic', tables,
                                                                  False, True, fetch=fetch),
        'features_result_rows': lambda fetch: eng.decode_features(th_d, to_d, 4, 4, 'bicubic', tables, fetch=fetch,
                                                                  frames=frames),
        'features_noisy_plane': lambda fetch: eng.decode_features(tn_d, to_d, 4, 4, 'bicubic', tables, fetch=fetch),
    }
    ref, ref_rows = {}, {}
    for k, f in calls.items():
        ref[k] = f(True)
        ref_rows[k] = eng.last_result_rows
    assert ref_rows['features_result_rows'] is not None and ref_rows['features_dev'] is None
    assert same(ref['features_dev'], ref['features_host']) and same(ref['features_dev'], ref['features_plan'])
    bad = {k: 0 for k in calls}
    order = list(calls)
    rng = np.random.RandomState(0)
    weights = np.array([1.0 if k != 'features_noisy_plane' else 0.05 for k in order])
    weights /= weights.sum()
    t0 = time.time()
    pending = []
    depth = 3
    for it in range(iters):
        k = order[rng.choice(len(order), p=weights)]
        calls[k](False)
        pending.append(k)
        if it % 97 == 0:
            depth = int(rng.randint(1, 13))
        while len(pending) >= depth:                # up to 12 calls in flight, mixed paths
            kk = pending.pop(0)
            if not (same(eng.fetch(n), ref[kk]) and same_rows(eng.last_result_rows, ref_rows[kk])):
                bad[kk] += 1
    while pending:
        kk = pending.pop(0)
        if not (same(eng.fetch(n), ref[kk]) and same_rows(eng.last_result_rows, ref_rows[kk])):
            bad[kk] += 1
    print(json.dumps({'iterations': iters, 'mismatches': bad, 'seconds': round(time.time() - t0, 1),
                      'persons': {k: sum(len(p) for p in v) for k, v in ref.items()},
                      'fused_redos': eng.fused_redo_count}))
    assert not any(bad.values())


if __name__ == '__main__':
    main()
