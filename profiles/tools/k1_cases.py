"""K1 (og_nms_topk_f32: nms_candidates + select) alone on inputs that do not flatter it:
BASELINE config 2 / 3 / 4 maps, synthetics with 30 % hot warp-rows, and thre <= 0 (every pixel
qualifies) on a dense noise floor and on a zero background.  GB/s = N*C*H*W*4 / time, against MEASURED_PEAKS hbm_gbs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402
import torch            # noqa: E402
from offsetguided_b200 import config as cfg          # noqa: E402
from offsetguided_b200.engine import DecoderEngine   # noqa: E402
from oracle import scenes                            # noqa: E402

if os.environ.get('OG_LIB'):             # a tuning variant of the library (offsetguided_b200.build(defines=..., out=...))
    import ctypes
    from offsetguided_b200 import _lib
    _v = ctypes.CDLL(os.environ['OG_LIB'])
    for _nm, (_rt, _at) in _lib.SIGNATURES.items():
        getattr(_v, _nm).restype, getattr(_v, _nm).argtypes = _rt, _at
    _lib._lib = _v

try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    PEAK = 6650.0


def timed(eng, heat, thre, reps=20):
    for _ in range(5):
        out = eng.nms_topk(heat, thre)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        out = eng.nms_topk(heat, thre)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    gbs = heat.numel() * 4 / (ms * 1e-3) / 1e9
    return ms, gbs, int((out[2] > 0).sum()), float(out[2].float().mean())


def report(name, eng, heat, thre, reps=20):
    ms, gbs, planes, cand = timed(eng, heat, thre, reps)
    print(json.dumps({'case': name, 'shape': list(heat.shape), 'thre': thre, 'ms': round(ms, 4), 'GBps': round(gbs, 1),
                      'frac_of_measured_peak': round(gbs / PEAK, 3), 'mean_candidates_kept_per_plane': round(cand, 2)}))


def main():
    coco, crowd = cfg.COCO_PERSON_SKELETON, cfg.CROWDPOSE_PERSON_SKELETON
    e17 = DecoderEngine(17, coco, topk=32, thre_hmp=0.04, dist_max=40, use_scale=True, person_thre=0.04)
    e14 = DecoderEngine(14, crowd, topk=64, thre_hmp=0.04, dist_max=40, use_scale=True, person_thre=0.04)
    h, _ = scenes.synth_hires_batch(1000, 8, 6, 640, 640, coco)
    heat = torch.from_numpy(h).cuda().repeat(8, 1, 1, 1).contiguous()
    report('cfg2: 64 x 17 x 640^2, 6 persons', e17, heat, 0.04)
    # broad blobs: about a third of the warp-rows (row x 128-column strip) hold values above the
    # threshold, with a realistic number of peaks (16 per plane)
    hb, _ = scenes.synth_hires_batch(5000, 8, 16, 640, 640, coco, sigma=12.0)
    frac = float((hb.reshape(8, 17, 640, 5, 128).max(axis=4) >= 0.04).mean())
    heat_hot = torch.from_numpy(hb).cuda().repeat(8, 1, 1, 1).contiguous()
    report('broad blobs, %.0f %% hot warp-rows: 64 x 17 x 640^2' % (100 * frac), e17, heat_hot, 0.04)
    del heat_hot
    # worst case for the candidate lists: 30 % of the warp-rows hold one ISOLATED value above the
    # threshold, i.e. ~960 peaks per plane (atomics on one counter per plane, O(n^2) ranking)
    rng = np.random.RandomState(3)
    hot = rng.uniform(0, 0.02, size=(8, 17, 640, 640)).astype(np.float32)
    mask = rng.uniform(size=(8, 17, 640, 5)) < 0.30
    cols = rng.randint(0, 128, size=(8, 17, 640, 5))
    n_i, c_i, r_i, s_i = np.nonzero(mask)
    hot[n_i, c_i, r_i, np.minimum(s_i * 128 + cols[n_i, c_i, r_i, s_i], 639)] = rng.uniform(0.05, 0.9, size=n_i.size)
    heat_hot = torch.from_numpy(hot).cuda().repeat(8, 1, 1, 1).contiguous()
    report('960 isolated peaks per plane (30 % hot warp-rows): 64 x 17 x 640^2', e17, heat_hot, 0.04)
    del heat_hot
    # thre <= 0 (the exact joint_dets API).  The rendered maps carry noise on every pixel (45 k positive
    # local maxima per plane): every plane overflows its candidate list and is selected by the
    # one-CTA-per-plane radix selection.  A map whose background is exactly zero (what hmp_NMS itself
    # returns) is listed by the streaming pass and completed from the zeros.
    report('thre = 0, dense noise floor (radix selection per plane): 64 x 17 x 640^2', e17, heat, 0.0, reps=5)
    report('thre = -1 (exact joint_dets), dense noise floor: 16 x 17 x 640^2', e17, heat[:16].contiguous(), -1.0, reps=3)
    sparse = torch.where(heat >= 0.02, heat, torch.zeros_like(heat))
    report('thre = 0, zero background: 64 x 17 x 640^2', e17, sparse, 0.0)
    report('thre = -1 (exact joint_dets), zero background: 64 x 17 x 640^2', e17, sparse, -1.0)
    del heat, sparse
    h, _ = scenes.synth_hires_batch(2000, 8, 20, 640, 640, crowd, n_channels=14)
    heat = torch.from_numpy(h).cuda().repeat(4, 1, 1, 1).contiguous()
    report('cfg3: 32 x 14 x 640^2, 20 persons, K = 64', e14, heat, 0.04)
    del heat
    h, _ = scenes.synth_hires_batch(3000, 4, 6, 1024, 1024, coco)
    heat = torch.from_numpy(h).cuda().repeat(16, 1, 1, 1).contiguous()
    report('cfg4: 64 x 17 x 1024^2, 6 persons', e17, heat, 0.04)


if __name__ == '__main__':
    main()
