"""K3 alone (og_group_f32 on the limb tables of the bench workloads, device outputs), CUDA events:
images/s per SM at 8 / 64 / 148 / 1184 images."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ctypes           # noqa: E402
import torch            # noqa: E402
from offsetguided_b200 import _lib, config as cfg    # noqa: E402
from offsetguided_b200.engine import DecoderEngine, _ptr, _stream_ptr   # noqa: E402
from oracle import scenes                            # noqa: E402


def limbs_of(persons, k, n_src=8):
    skel = cfg.COCO_PERSON_SKELETON
    h, o = scenes.synth_hires_batch(1000 + persons, n_src, persons, 640, 640, skel)
    eng = DecoderEngine(17, skel, topk=k, thre_hmp=0.04, dist_max=40, use_scale=True, person_thre=0.04)
    eng.decode_maps(torch.from_numpy(h).cuda(), torch.from_numpy(o).cuda())
    return eng, eng.last_intermediates(n_src)[2]


def main():
    for persons, k in ((6, 32), (20, 64)):
        eng, lb = limbs_of(persons, k)
        for n in (1, 8, 64, 148, 1184):
            limbs = lb.repeat((n + 7) // 8, 1, 1, 1)[:n].contiguous()
            cap = n * 19 * k
            poses = torch.empty((cap, 17, 6), dtype=torch.float32, device='cuda')
            meta = torch.empty((2 * n + 1,), dtype=torch.int32, device='cuda')
            def run():
                _lib.check(eng.lib.og_group_f32(eng._h, _ptr(limbs), n, _ptr(poses), cap, _ptr(meta[:n]),
                                                _ptr(meta[n:2 * n]), _ptr(meta[2 * n:]), _stream_ptr(eng.device)))
            for _ in range(5):
                run()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(20):
                run()
            b.record()
            torch.cuda.synchronize()
            us = 1e3 * a.elapsed_time(b) / 20
            print(json.dumps({'persons': persons, 'K': k, 'images': n, 'us_per_launch_incl_prepare': round(us, 2),
                              'images_per_s': round(n / (us * 1e-6)), 'images_per_s_per_SM': round(n / (us * 1e-6) / min(n, 148)),
                              'total_persons': int(meta[2 * n])}))


if __name__ == '__main__':
    main()
