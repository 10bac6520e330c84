"""Loaders for the committed golden fixtures (tests/golden/*.npz, written by
tests/golden/make_golden.py from the reference's own outputs)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def add_noise(x, seed, amp):
    """Same formula as make_golden.add_noise (fixtures store seed + amplitude)."""
    if amp <= 0:
        return x
    return x + np.random.RandomState(int(seed)).uniform(0, float(amp), size=x.shape).astype(np.float32)


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    return {k: z[k] for k in z.files}


def load_limbs_case(name):
    d = load(name)
    d['heat'] = add_noise(d['heat'], d['noise_seed'], d['noise_amp'])
    d['skeleton'] = [tuple(int(v) for v in p) for p in d['skeleton']]
    for k in ('topk', 'n_keypoints'):
        d[k] = int(d[k])
    for k in ('thre_hmp', 'person_thre', 'dist_max', 'min_len'):
        d[k] = float(d[k])
    d['resize_mode'] = str(d['resize_mode']) if 'resize_mode' in d else 'bicubic'
    return d


def load_poses_case(name):
    d = load(name)
    d['hmp'] = add_noise(d['hmp'], d['noise_seed'], d['noise_amp'])
    d['flip_test'] = bool(d['flip_test'])
    d['topk'] = int(d['topk'])
    for k in ('thre_hmp', 'person_thre', 'dist_max', 'min_len'):
        d[k] = float(d[k])
    d['resize_mode'] = str(d['resize_mode']) if 'resize_mode' in d else 'bicubic'
    return d


OPTIONAL_VARIANTS = {
    # name: (include_scale, include_jitter_offset, use_jitter_offset, flip_test, cat_flip_offs)
    'scale_jitter_flip': (True, True, True, True, False),
    'jitter_noflip': (False, True, True, False, False),
    'jitter_unused': (False, True, False, False, False),
    'cat_flip': (False, False, True, True, True),
    'scale_cat_flip': (True, False, True, True, True),
}


def load_optional_heads():
    """Inputs (incl. the seeded scale / jitter maps) and the reference's poses per variant."""
    d = load('poses_optional_heads')
    n2, _, h, w = d['hmp'].shape
    rng = np.random.RandomState(int(d['noise_seed']))
    d['scm'] = rng.uniform(2, 60, size=(n2, 17, h, w)).astype(np.float32)
    d['jom'] = rng.uniform(-1.5, 1.5, size=(n2, 2, h, w)).astype(np.float32)
    return d


def load_group_fuzz():
    d = load('group_fuzz')
    cases = []
    for i in range(int(d['n'])):
        c, thre, sort_dim, use_scale = d[f'meta_{i}']
        cases.append(dict(limbs=d[f'limbs_{i}'],
                          skeleton=[tuple(int(v) for v in p) for p in d[f'skel_{i}']],
                          n_keypoints=int(c), person_thre=float(thre), sort_dim=int(sort_dim),
                          use_scale=bool(use_scale), poses=d[f'poses_{i}']))
    return cases


def split_poses(poses, counts):
    out, at = [], 0
    for c in counts:
        out.append(poses[at:at + int(c)])
        at += int(c)
    return out


def tolerances(name, rtol):
    """(limb rtol, min_dist atol, pose rtol) of a poses fixture: the same for all of them.
    (All fixtures are generated from C-contiguous NCHW tensors, the layout a network emits: ATen's
    bilinear resize has a separate channels-last kernel whose results differ by 1 ulp, which an
    earlier generator — np.stack of the encoder's (H, W, C) arrays keeps that memory order — ran
    into; tests/golden/make_golden.py explains.)"""
    return rtol, 1e-6, rtol


def compare_limbs(got, ref, thre_hmp, rtol=1e-5, dist_atol=1e-6):
    """Compare two (N, L, K, 13) limb tables on the rows that can influence
    grouping: rows whose from-candidate is above threshold.  Integer-valued
    columns (x, y, ids) must be equal; float columns agree within rtol.
    If the matched to-candidate is below threshold the row is inert (min_dist
    ~1e5 fails every distance gate) and only that property is checked."""
    assert got.shape == ref.shape, (got.shape, ref.shape)
    live = ref[..., 2] >= np.float32(thre_hmp)
    assert np.array_equal(live, got[..., 2] >= np.float32(thre_hmp)), 'live from-candidates differ'
    to_live = live & (ref[..., 5] >= np.float32(thre_hmp))
    inert = live & ~to_live
    assert np.all(got[inert][:, 8] > 5e4) and np.all(ref[inert][:, 8] > 5e4)
    g, r = got[to_live], ref[to_live]
    for col in (0, 1, 3, 4, 6, 7, 11, 12):
        assert np.array_equal(g[:, col], r[:, col]), f'limb column {col} differs'
    for col in (2, 5, 8, 9, 10):
        np.testing.assert_allclose(g[:, col], r[:, col], rtol=rtol, atol=dist_atol if col == 8 else 1e-6,
                                   err_msg=f'limb column {col}')
    return int(to_live.sum())


def compare_poses(got, ref, rtol=1e-5, exact=False):
    """Two (M, C, 6) pose arrays: same person count and order, x / y / id exact,
    v / s / limb score within rtol (bit-exact when ``exact``)."""
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if exact:
        assert np.array_equal(got, ref)
        return
    for col in (0, 1, 5):
        assert np.array_equal(got[..., col], ref[..., col]), f'pose column {col} differs'
    for col in (2, 3, 4):
        np.testing.assert_allclose(got[..., col], ref[..., col], rtol=rtol, atol=1e-7)


def compare_dets_tie_aware(got_s, got_i, ref_s, ref_i, thre):
    """Top-K tables (N, C, K) against the reference's own torch.topk output, TIE-AWARE: within a
    plane the live (score >= thre) scores must be equal slot by slot (both are sorted descending),
    and for every distinct score the SET of indices must be equal — the order of equal values is
    library-defined in torch.topk.  Returns the number of tied groups seen."""
    assert got_s.shape == ref_s.shape and got_i.shape == ref_i.shape
    groups = 0
    n, c, _ = ref_s.shape
    for i in range(n):
        for j in range(c):
            live = ref_s[i, j] >= np.float32(thre)
            assert np.array_equal(got_s[i, j] >= np.float32(thre), live)
            assert np.array_equal(got_s[i, j][live], ref_s[i, j][live]), 'live scores differ'
            for v in np.unique(ref_s[i, j][live]):
                a = np.sort(got_i[i, j][live & (got_s[i, j] == v)])
                b = np.sort(ref_i[i, j][live & (ref_s[i, j] == v)])
                assert np.array_equal(a, b), 'index sets of a tied group differ'
                groups += int(len(b) > 1)
    return groups
