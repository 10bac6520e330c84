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
