"""Per-phase cycle profile of the K3 grouping kernel (library built with -DOG_K3_PROFILE into
build/k3_variants/prof.so: python -c "from offsetguided_b200 import build; build.build(defines={'OG_K3_PROFILE': 1}, out='build/k3_variants/prof.so')").
Default: the one-warp-per-image kernel; OG_K3_WARP_ROWS=0 profiles the CTA kernel."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from offsetguided_b200 import _lib, engine      # noqa: E402
from offsetguided_b200 import config as cfg     # noqa: E402
from oracle import scenes                       # noqa: E402

NAMES_CTA = ['top (wait rows, flags)', 'stage kept rows', 'match pairs', 'apply', 'merge search',
             'merge apply + new scan', 'init new rows', 'final (score/sort)', 'output', '-']
NAMES_WARP = ['top (wait rows, prefetch)', 'match pairs', 'apply', '-', 'merge search',
              'merge apply', 'new persons', 'final (score/sort)', 'output', '-']
NAMES = NAMES_CTA if os.environ.get('OG_K3_WARP_ROWS') == '0' else NAMES_WARP


def main():
    lib = ctypes.CDLL(os.path.join(ROOT, 'build', 'k3_variants', 'prof.so'))
    for nm, (restype, argtypes) in _lib.SIGNATURES.items():
        fn = getattr(lib, nm)
        fn.restype, fn.argtypes = restype, argtypes
    _lib._lib = lib
    skel = cfg.COCO_PERSON_SKELETON
    for persons, k, nimg in ((6, 32, 64), (20, 64, 32)):
        h1, o1 = scenes.synth_hires_batch(1000 + persons, 8, persons, 640, 640, skel)
        heat = torch.from_numpy(h1).cuda().repeat(nimg // 8, 1, 1, 1).contiguous()
        offs = torch.from_numpy(o1).cuda().repeat(nimg // 8, 1, 1, 1).contiguous()
        eng = engine.DecoderEngine(17, skel, topk=k, thre_hmp=0.04, dist_max=40, use_scale=True, person_thre=0.04)
        eng.enable_stage_timing(True)
        for _ in range(3):
            eng.decode_maps(heat, offs)
        out = (ctypes.c_uint64 * 16)()
        _lib.check(lib.og_debug_k3_profile(out, 1))
        reps = 10
        k3 = []
        for _ in range(reps):
            eng.decode_maps(heat, offs)
            k3.append(eng.last_stage_times_ms()['k3'])
        _lib.check(lib.og_debug_k3_profile(out, 1))
        ctas = out[15]
        print('workload: %d persons, K=%d, %d images; k3 = %.1f us; CTAs profiled %d' %
              (persons, k, nimg, 1e3 * np.mean(k3), ctas))
        tot = sum(out[i] for i in range(10))
        for i, nm in enumerate(NAMES):
            print('  %-22s %9.0f cycles/image  %5.1f%%' % (nm, out[i] / ctas, 100.0 * out[i] / tot))
        print('  %-22s %9.0f cycles/image' % ('total', tot / ctas))
        eng.close()


if __name__ == '__main__':
    main()
