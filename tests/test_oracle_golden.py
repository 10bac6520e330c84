"""CPU suite: the oracle (oracle/ref_oracle.py) against the committed outputs of the
reference itself (tests/golden/*.npz, written by tests/golden/make_golden.py)."""
import numpy as np
import pytest

import golden_io as gio
from oracle import ref_oracle as ro
from oracle import scenes
from offsetguided_b200 import config as cfg


@pytest.mark.parametrize('name', ['limbs_coco_a', 'limbs_coco_noise', 'limbs_crowdpose'])
def test_limbs_and_poses_match_reference(name):
    d = gio.load_limbs_case(name)
    limbs, dets = ro.generate_limbs(d['heat'], d['offs'], d['skeleton'], d['topk'], d['thre_hmp'],
                                    d['min_len'], 1, 1, return_dets=True)
    live = d['det_scores'] >= np.float32(d['thre_hmp'])
    assert live.sum() > 50
    assert np.array_equal(dets[0][live], d['det_scores'][live])       # bit-exact peak values
    assert np.array_equal(dets[1][live], d['det_inds'][live])         # bit-exact peak indices
    rows = gio.compare_limbs(limbs, d['limbs'], d['thre_hmp'], rtol=1e-6)
    assert rows > 100
    # min_dist / len use the fma formula probed from ATen: bit-exact columns
    both = (d['limbs'][..., 2] >= np.float32(d['thre_hmp'])) & (d['limbs'][..., 5] >= np.float32(d['thre_hmp']))
    assert np.array_equal(limbs[both][:, 8], d['limbs'][both][:, 8])
    assert np.array_equal(limbs[both][:, 9], d['limbs'][both][:, 9])
    # grouping on the REFERENCE's limbs must be bit-exact
    ref_poses = gio.split_poses(d['poses'], d['pose_counts'])
    for img, ref in enumerate(ref_poses):
        got = ro.group_skeletons(d['limbs'][img], d['skeleton'], d['n_keypoints'], d['person_thre'],
                                 2, d['dist_max'], True)
        gio.compare_poses(got, ref, exact=True)
        got2 = ro.group_skeletons(limbs[img], d['skeleton'], d['n_keypoints'], d['person_thre'],
                                  2, d['dist_max'], True)
        gio.compare_poses(got2, ref, rtol=1e-6)


def test_group_fuzz_bit_exact():
    cases = gio.load_group_fuzz()
    assert len(cases) >= 100
    stats = {}
    for c in cases:
        got = ro.group_skeletons(c['limbs'], c['skeleton'], c['n_keypoints'], c['person_thre'],
                                 c['sort_dim'], 40, c['use_scale'], stats)
        gio.compare_poses(got, c['poses'], exact=True)
    # the fixture exercises every branch of the reference's grouping
    for key in ('case2', 'case1', 'dup_person_write', 'merge', 'share3', 'cancel_new'):
        assert stats[key] > 0, key


@pytest.mark.parametrize('name', ['poses_cfg1', 'poses_cfg2_flip', 'poses_inf_background',
                                  'poses_inf_background_flip', 'poses_bilinear_flip'])
def test_generate_poses_matches_reference(name):
    d = gio.load_poses_case(name)
    fl, rs = cfg.offset_hflip(cfg.COCO_KEYPOINTS, cfg.COCO_PERSON_SKELETON)
    poses, limbs = ro.generate_poses(
        d['hmp'], d['omp'], cfg.COCO_PERSON_SKELETON, 17, topk=d['topk'], thre_hmp=d['thre_hmp'],
        min_len=d['min_len'], person_thre=d['person_thre'], dist_max=d['dist_max'], use_scale=True,
        flip_test=d['flip_test'], kp_flips=cfg.heatmap_hflip(cfg.COCO_KEYPOINTS), limb_flips=fl,
        limb_reserve=rs, return_limbs=True, resize_mode=d['resize_mode'])
    lr, da, pr = gio.tolerances(name, 1e-6)
    assert gio.compare_limbs(limbs, d['limbs'], d['thre_hmp'], rtol=lr, dist_atol=da) > 50
    ref = gio.split_poses(d['poses'], d['pose_counts'])
    assert len(poses) == len(ref)
    for p, r in zip(poses, ref):
        gio.compare_poses(p, r, rtol=pr)


def test_resize_bit_exact_against_aten():
    d = gio.load('resize_small')
    assert np.array_equal(ro.resize(d['x'], 4, 'bicubic'), d['bicubic4'])
    assert np.array_equal(ro.resize(d['x'], 4, 'bilinear'), d['bilinear4'])
    assert np.array_equal(ro.resize(d['x2'], 2, 'bicubic'), d['bicubic2'])
    assert np.array_equal(ro.resize(d['x2'], 2, 'bilinear'), d['bilinear2'])
    assert np.array_equal(ro.resize(d['x8'], 8, 'bicubic'), d['bicubic8'])
    assert np.array_equal(ro.resize(d['x8'], 8, 'bilinear'), d['bilinear8'])


def test_resize_probes_of_pose_fixtures():
    d = gio.load_poses_case('poses_cfg1')
    hr = ro.resize(d['hmp'], 4, 'bicubic')
    assert np.array_equal(hr[:, :, ::37, ::41], d['heat_hr_probe'])
    orr = ro.resize(d['omp'], 4, 'bilinear')
    assert np.array_equal(orr[:, :, ::37, ::41], d['offs_hr_probe'])


def test_nms_semantics_edge_cases():
    h = np.zeros((1, 1, 5, 6), np.float32)
    h[0, 0, 0, 0] = -1.0           # negative border value: suppressed by the zero padding
    h[0, 0, 2, 2] = h[0, 0, 2, 3] = 0.7   # plateau: both survive
    h[0, 0, 4, 5] = 0.3
    out = ro.hmp_nms(h)
    assert out[0, 0, 0, 0] == 0
    assert out[0, 0, 2, 2] == np.float32(0.7) and out[0, 0, 2, 3] == np.float32(0.7)
    assert out[0, 0, 4, 5] == np.float32(0.3)
    s, i, ys, xs = ro.topk_channel(out, 4)
    assert list(i[0, 0][:3]) == [14, 15, 29]      # ties ordered by index
    assert list(ys[0, 0][:3]) == [2, 2, 4] and list(xs[0, 0][:3]) == [2, 3, 5]


def test_numpy_pairwise_sum_restated():
    """K3 restates numpy's float32 pairwise summation; check the restatement."""
    rng = np.random.RandomState(3)

    def pw(a):
        n = len(a)
        f = np.float32
        if n < 8:
            s = f(0)
            for x in a:
                s = f(s + x)
            return s
        r = [f(x) for x in a[:8]]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = f(r[j] + a[i + j])
            i += 8
        s = f(f(f(r[0] + r[1]) + f(r[2] + r[3])) + f(f(r[4] + r[5]) + f(r[6] + r[7])))
        while i < n:
            s = f(s + a[i])
            i += 1
        return s
    for _ in range(3000):
        a = rng.uniform(0, 1, size=rng.randint(1, 65)).astype(np.float32)
        assert pw(a) == a.sum()


def test_scene_renderer_is_deterministic():
    a = scenes.render_batch(11, 1, 4, 256, 192, cfg.COCO_PERSON_SKELETON)
    b = scenes.render_batch(11, 1, 4, 256, 192, cfg.COCO_PERSON_SKELETON)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[0].shape == (1, 17, 48, 64) and a[1].shape == (1, 38, 48, 64)
    assert a[0].max() > 0.9


@pytest.mark.parametrize('name', list(gio.OPTIONAL_VARIANTS))
def test_optional_heads_match_reference(name):
    """Keypoint-scale maps, jitter-offset maps (used / unused) and cat_flip_offs."""
    d = gio.load_optional_heads()
    inc_scale, inc_jit, use_jit, flip, cat = gio.OPTIONAL_VARIANTS[name]
    n = d['hmp'].shape[0] // 2
    sel = slice(None) if flip else slice(0, n)
    fl, rs = cfg.offset_hflip(cfg.COCO_KEYPOINTS, cfg.COCO_PERSON_SKELETON)
    poses = ro.generate_poses(
        d['hmp'][sel], d['omp'][sel], cfg.COCO_PERSON_SKELETON, 17, topk=16, thre_hmp=0.06, min_len=0.5,
        person_thre=0.06, dist_max=40, use_scale=True, flip_test=flip,
        kp_flips=cfg.heatmap_hflip(cfg.COCO_KEYPOINTS), limb_flips=fl, limb_reserve=rs,
        scmps=d['scm'][sel] if inc_scale else None, jomps=d['jom'][sel] if inc_jit else None,
        use_jitter=use_jit, cat_flip_offs=cat)
    ref = gio.split_poses(d[name + '_poses'], d[name + '_counts'])
    assert len(poses) == len(ref) and sum(len(p) for p in ref) >= 6
    for p, r in zip(poses, ref):
        gio.compare_poses(p, r, rtol=1e-6) if not inc_jit or not use_jit else None
        assert p.shape == r.shape and np.array_equal(p[..., 5], r[..., 5])
        np.testing.assert_allclose(p, r, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('name', ['poses_cfg1', 'poses_cfg2_flip', 'poses_bilinear_flip'])
def test_torch_eager_baseline_matches_reference_fixtures(name):
    """oracle/torch_eager.py (bench.py's eager-PyTorch baseline, an independent formulation of the
    path with stock ATen operators) decodes the reference fixtures to the reference's poses."""
    import torch
    from oracle import torch_eager as te
    d = gio.load_poses_case(name)
    fl, rs = cfg.offset_hflip(cfg.COCO_KEYPOINTS, cfg.COCO_PERSON_SKELETON)
    got = te.generate_poses(torch.from_numpy(d['hmp']), torch.from_numpy(d['omp']), cfg.COCO_PERSON_SKELETON, 17,
                            topk=d['topk'], thre_hmp=d['thre_hmp'], min_len=d['min_len'],
                            person_thre=d['person_thre'], dist_max=d['dist_max'], use_scale=True, stride=4,
                            resize_mode=d['resize_mode'], flip_test=d['flip_test'],
                            kp_flips=cfg.heatmap_hflip(cfg.COCO_KEYPOINTS), limb_flips=fl, limb_reserve=rs)
    ref = gio.split_poses(d['poses'], d['pose_counts'])
    assert len(got) == len(ref)
    for p, r in zip(got, ref):
        gio.compare_poses(p, r, rtol=1e-6)


def test_scored_offset_matches_reference_fixture():
    """oracle scored_offset == the reference's decoder.scored_offset output, bit for bit
    (decoder/offset.py:8-43; k = 3 is the call site's window, 7 the default)."""
    d = gio.load('scored_offset')
    jf, jt = ro.pack_jtypes(cfg.COCO_PERSON_SKELETON)
    for tag in ('a', 'b'):
        for ks in (3, 7):
            got = ro.scored_offset(d['hmp_' + tag], d['off_' + tag], jf, jt, ks)
            assert np.array_equal(got, d['out_%s_k%d' % (tag, ks)])


def test_generate_poses_scored_off_matches_reference():
    """generate_poses(flip_test=True, scored_off=True) of the reference on the config-2 inputs."""
    d = gio.load_poses_case('poses_cfg2_flip')
    s = gio.load('poses_scored_off')
    kp = cfg.heatmap_hflip(cfg.COCO_KEYPOINTS)
    fl, rs = cfg.offset_hflip(cfg.COCO_KEYPOINTS, cfg.COCO_PERSON_SKELETON)
    fh, fo = ro.flip_augment(d['hmp'], d['omp'], kp, fl, rs)
    jf, jt = ro.pack_jtypes(cfg.COCO_PERSON_SKELETON)
    fo = ro.scored_offset(fh, fo, jf, jt, 3)
    got = ro.generate_poses(fh, fo, cfg.COCO_PERSON_SKELETON, 17, topk=32, thre_hmp=0.04, min_len=0.5,
                            person_thre=0.04, dist_max=40, use_scale=True)
    ref = gio.split_poses(s['poses'], s['pose_counts'])
    assert len(got) == len(ref) == 2
    for p, r in zip(got, ref):
        gio.compare_poses(p, r, rtol=1e-5)


def test_tied_peaks_fixture_tie_aware():
    """Scenes whose bicubic x4 heat maps hold equal-valued above-threshold peaks (2-pixel plateaus,
    SURVEY 8c): dets compared with the reference's torch.topk output as sets within tied groups,
    final poses exactly (the generator checked that they do not depend on the order)."""
    d = gio.load('poses_tied_peaks')
    assert int(d['ties']) >= 2 and bool(d['poses_order_independent'])
    thre = float(d['thre_hmp'])
    k, pt, dm = int(d['topk']), float(d['person_thre']), float(d['dist_max'])
    hr = ro.resize(d['hmp'], 4, 'bicubic')
    limbs, dets = ro.generate_limbs(hr, ro.resize(d['omp'], 4, 'bilinear'), cfg.COCO_PERSON_SKELETON, k, thre,
                                    0.5, 4, 4, return_dets=True)
    poses = [ro.group_skeletons(l, cfg.COCO_PERSON_SKELETON, 17, pt, 2, dm, True) for l in limbs]
    assert gio.compare_dets_tie_aware(dets[0], dets[1], d['det_scores'], d['det_inds'], thre) >= 2
    for p, r in zip(poses, gio.split_poses(d['poses'], d['pose_counts'])):
        gio.compare_poses(p, r, rtol=1e-5)
