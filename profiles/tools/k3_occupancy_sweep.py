"""K3 throughput against the batch size and the number of person-table rows kept in shared
memory (OG_K3_ROWS): with fewer rows several images share an SM.  Times og_group_f32 alone
(prepare + grouping kernels) on limb tables of the bench workload, tiled to N images."""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    from offsetguided_b200 import _lib
    from offsetguided_b200 import config as cfg
    from offsetguided_b200.engine import DecoderEngine, _ptr, _stream_ptr
    from oracle import scenes
    skel = cfg.COCO_PERSON_SKELETON
    persons = int(os.environ.get('K3_PERSONS', '6'))
    heat, offs = scenes.synth_hires_batch(1000, 8, persons, 640, 640, skel)
    eng = DecoderEngine(17, skel, topk=32, thre_hmp=0.04, dist_max=40, use_scale=True, person_thre=0.04)
    eng.decode_maps(torch.from_numpy(heat).cuda(), torch.from_numpy(offs).cuda())
    _, _, limbs8 = eng.last_intermediates(8)
    dev = eng.device
    for n in (64, 148, 296, 592, 1184):
        limbs = limbs8.repeat((n + 7) // 8, 1, 1, 1)[:n].contiguous()
        cap = n * 19 * 32
        poses = torch.empty((cap, 17, 6), dtype=torch.float32, device=dev)
        meta = torch.empty((2 * n + 1,), dtype=torch.int32, device=dev)

        def run():
            _lib.check(eng.lib.og_group_f32(eng._h, _ptr(limbs), n, _ptr(poses), cap, _ptr(meta[:n]),
                                            _ptr(meta[n:2 * n]), _ptr(meta[2 * n:]), _stream_ptr(dev)))
        for _ in range(3):
            run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            run()
        b.record()
        torch.cuda.synchronize()
        us = 1e3 * a.elapsed_time(b) / 20
        total = int(meta[2 * n].item())
        print(json.dumps({'rows': os.environ.get('OG_K3_ROWS', 'default'), 'persons_per_image': persons,
                          'images': n, 'k3_us': round(us, 1), 'images_per_s': round(n / us * 1e6),
                          'images_per_s_per_sm': round(n / us * 1e6 / min(n, 148)), 'persons': total}))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        child()
    else:
        for persons in ('6', '20'):
            for rows in ('256', '128', '96', '64'):
                env = dict(os.environ, OG_K3_ROWS=rows, K3_PERSONS=persons)
                subprocess.run([sys.executable, os.path.abspath(__file__), 'child'], env=env, check=False)
