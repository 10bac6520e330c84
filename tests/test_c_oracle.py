"""CPU suite: the C oracle (oracle/og_oracle.c) against the committed reference outputs
and against the numpy oracle."""
import numpy as np
import pytest

import golden_io as gio
from oracle import c_oracle as co
from oracle import ref_oracle as ro
from offsetguided_b200 import config as cfg


def test_c_resize_bit_exact_against_aten():
    d = gio.load('resize_small')
    for key, xk, scale, mode in (('bicubic4', 'x', 4, 'bicubic'), ('bilinear4', 'x', 4, 'bilinear'),
                                 ('bicubic2', 'x2', 2, 'bicubic'), ('bilinear2', 'x2', 2, 'bilinear'),
                                 ('bicubic8', 'x8', 8, 'bicubic'), ('bilinear8', 'x8', 8, 'bilinear')):
        assert np.array_equal(co.resize(d[xk], scale, mode), d[key]), key


@pytest.mark.parametrize('name', ['limbs_coco_a', 'limbs_coco_noise', 'limbs_crowdpose'])
def test_c_limbs_and_groups_match_reference(name):
    d = gio.load_limbs_case(name)
    limbs, dets = co.generate_limbs(d['heat'], d['offs'], d['skeleton'], d['topk'], d['thre_hmp'],
                                    d['min_len'], 1, 1, return_dets=True)
    nd = ro.joint_dets(d['heat'], d['topk'])
    assert np.array_equal(dets[0], nd[0]) and np.array_equal(dets[1], nd[1])     # incl. zero filler
    assert gio.compare_limbs(limbs, d['limbs'], d['thre_hmp'], rtol=1e-6) > 100
    poses = co.group_batch(d['limbs'], d['skeleton'], d['n_keypoints'], d['person_thre'], 2,
                           d['dist_max'], True)
    for p, r in zip(poses, gio.split_poses(d['poses'], d['pose_counts'])):
        gio.compare_poses(p, r, exact=True)


def test_c_group_fuzz_bit_exact():
    for c in gio.load_group_fuzz():
        got = co.group_skeletons(c['limbs'], c['skeleton'], c['n_keypoints'], c['person_thre'],
                                 c['sort_dim'], 40, c['use_scale'])
        gio.compare_poses(got, c['poses'], exact=True)


@pytest.mark.parametrize('name', ['poses_cfg1', 'poses_cfg2_flip', 'poses_inf_background',
                                  'poses_inf_background_flip', 'poses_bilinear_flip'])
def test_c_generate_poses_matches_reference(name):
    d = gio.load_poses_case(name)
    fl, rs = cfg.offset_hflip(cfg.COCO_KEYPOINTS, cfg.COCO_PERSON_SKELETON)
    poses, limbs = co.generate_poses(
        d['hmp'], d['omp'], cfg.COCO_PERSON_SKELETON, 17, topk=d['topk'], thre_hmp=d['thre_hmp'],
        min_len=d['min_len'], person_thre=d['person_thre'], dist_max=d['dist_max'], use_scale=True,
        flip_test=d['flip_test'], kp_flips=cfg.heatmap_hflip(cfg.COCO_KEYPOINTS), limb_flips=fl,
        limb_reserve=rs, return_limbs=True, resize_mode=d['resize_mode'])
    lr, da, pr = gio.tolerances(name, 1e-6)
    assert gio.compare_limbs(limbs, d['limbs'], d['thre_hmp'], rtol=lr, dist_atol=da) > 50
    for p, r in zip(poses, gio.split_poses(d['poses'], d['pose_counts'])):
        gio.compare_poses(p, r, rtol=pr)
    # heat-map values after flip fusion + bicubic x4 are bit-identical to the reference's
    live = d['det_scores'] >= np.float32(d['thre_hmp'])
    hm, om = (co.flip_augment(d['hmp'], d['omp'], cfg.heatmap_hflip(cfg.COCO_KEYPOINTS), fl, rs)
              if d['flip_test'] else (d['hmp'], d['omp']))
    dets = co.joint_dets(co.resize(hm, 4, d['resize_mode']), d['topk'])
    assert np.array_equal(dets[0][live], d['det_scores'][live])
    assert np.array_equal(dets[1][live], d['det_inds'][live])


def test_c_oracle_random_groups_equal_numpy_oracle():
    rng = np.random.RandomState(5)
    skel = cfg.COCO_PERSON_SKELETON
    for case in range(30):
        k = int(rng.choice([4, 8, 16]))
        pool = int(rng.choice([2, 4, 9]))
        limbs = np.zeros((19, k, 13), np.float32)
        xy = rng.randint(1, 600, size=(17, pool, 2)).astype(np.float32)
        ids = rng.randint(0, 640 * 640, size=(17, pool))
        for l, (jf, jt) in enumerate(skel):
            sc = (rng.permutation(k) + rng.uniform(0.1, 0.9, size=k)).astype(np.float32) / k
            for r in range(k):
                a, b = rng.randint(pool), rng.randint(pool)
                limbs[l, r] = (xy[jf, a, 0], xy[jf, a, 1], 0.5, xy[jt, b, 0], xy[jt, b, 1], 0.6,
                               ids[jf, a] + jf * 409600, ids[jt, b] + jt * 409600,
                               rng.uniform(0, 50), 10, sc[r], 4, 4)
        a = co.group_skeletons(limbs, skel, 17, 0.06, 2, 40, True)
        b = ro.group_skeletons(limbs, skel, 17, 0.06, 2, 40, True)
        assert a.shape == b.shape and np.array_equal(a, b), case


def test_c_oracle_random_maps_equal_numpy_oracle():
    """Random full decode (flip fusion, x2 / x4 resize in both modes, NMS / top-K with plateaus and
    negative borders, limbs, grouping) through both oracles: the multi-threaded C restatement that
    checks the large GPU cases must agree with the numpy restatement it was derived from."""
    rng = np.random.RandomState(11)
    skel = [(0, 1), (1, 2), (0, 2), (2, 3)]
    kp_flips, limb_flips, limb_reserve = [0, 2, 1, 3], [2, 1, 0, 3], [1]
    for case in range(6):
        stride = int(rng.choice([2, 4]))
        mode = str(rng.choice(['bicubic', 'bilinear']))
        flip = bool(case % 2)
        n, h, w = 2, int(rng.randint(20, 30)), int(rng.randint(40, 52))      # output H + W > 128
        hmp = np.zeros((2 * n if flip else n, 4, h, w), np.float32)
        for img in range(hmp.shape[0]):
            for c in range(4):
                for _ in range(5):
                    y, x = rng.randint(0, h), rng.randint(0, w)
                    hmp[img, c, y, x] = rng.uniform(0.2, 1.0)
                hmp[img, c, 3, 5:7] = 0.5                                   # plateau
        hmp[:, :, 0, :] -= rng.uniform(0, 0.3, size=(hmp.shape[0], 4, w)).astype(np.float32)
        omp = rng.uniform(-12, 12, size=(hmp.shape[0], 8, h, w)).astype(np.float32)
        kw = dict(topk=8, thre_hmp=0.1, min_len=0.5, person_thre=0.1, dist_max=30.0, use_scale=True,
                  hmp_stride=stride, off_stride=stride, resize_mode=mode, flip_test=flip,
                  kp_flips=kp_flips, limb_flips=limb_flips, limb_reserve=limb_reserve, return_limbs=True)
        pa, la = co.generate_poses(hmp, omp, skel, 4, **kw)
        pb, lb = ro.generate_poses(hmp, omp, skel, 4, **kw)
        assert gio.compare_limbs(la, lb, 0.1, rtol=1e-6) >= 4, case
        assert len(pa) == len(pb) and sum(len(p) for p in pb) >= 2
        for a, b in zip(pa, pb):
            gio.compare_poses(a, b, rtol=1e-6)
