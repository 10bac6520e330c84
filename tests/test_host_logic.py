"""CPU suite: host-side logic of the drop-in package and the C-ABI surface."""
import argparse
import ctypes
import os
import re

import numpy as np
import pytest

from offsetguided_b200 import config as cfg
from offsetguided_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flip_tables_match_reference_probe():
    # values probed from the reference
    assert cfg.heatmap_hflip(cfg.COCO_KEYPOINTS) == [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]
    flips, reserve = cfg.offset_hflip(cfg.COCO_KEYPOINTS, cfg.COCO_PERSON_SKELETON)
    assert flips == [1, 0, 2, 4, 3, 5, 7, 6, 10, 11, 8, 9, 13, 12, 14, 17, 18, 15, 16]
    assert reserve == [2, 5, 14]


def test_skeleton_tables():
    assert len(cfg.COCO_PERSON_SKELETON) == 19
    assert len(cfg.KINEMATIC_TREE_SKELETON) == 16
    assert len(cfg.COCO_PERSON_WITH_REDUNDANT_SKELETON) == 31
    assert len(cfg.DENSER_COCO_PERSON_SKELETON) == 44
    assert len(cfg.REDUNDANT_CONNECTIONS) == 29
    for a, b in cfg.CROWDPOSE_PERSON_SKELETON:
        assert 0 <= a < 14 and 0 <= b < 14 and a != b
    fl, rs = cfg.offset_hflip(cfg.CROWDPOSE_KEYPOINTS, cfg.CROWDPOSE_PERSON_SKELETON)
    assert sorted(fl) == list(range(len(cfg.CROWDPOSE_PERSON_SKELETON)))


def _args(**over):
    from offsetguided_b200 import decoder
    p = argparse.ArgumentParser()
    decoder.decoder_cli(p)
    a = p.parse_args([])
    a.headnets = ['hmp', 'omp']
    a.strides = [4, 4]
    a.batch_size = 8
    a.include_scale = False
    a.include_jitter_offset = False
    for k, v in over.items():
        setattr(a, k, v)
    return a


def test_cli_defaults_match_reference():
    a = _args()
    assert (a.resize_mode, a.topk, a.thre_hmp, a.min_len, a.feat_stage) == ('bicubic', 48, 0.06, 0.5, -1)
    assert (a.person_thre, a.sort_dim, a.dist_max, a.use_scale, a.use_jitter_offset) == (0.06, 2, 20, True, True)
    p = argparse.ArgumentParser()
    from offsetguided_b200 import decoder
    decoder.decoder_cli(p)
    b = p.parse_args('--topk 32 --thre-hmp 0.04 --person-thre 0.04 --dist-max 40 --use-scale False'.split())
    assert b.topk == 32 and b.use_scale is False and b.dist_max == 40.0
    with pytest.raises(SystemExit):
        p.parse_args(['--use-scale', 'maybe'])


def test_decoder_factory_builds_without_gpu():
    from offsetguided_b200 import decoder
    pp = decoder.decoder_factory(_args(topk=32, dist_max=40))
    assert isinstance(pp, decoder.PostProcess)
    assert pp.skeleton == cfg.COCO_PERSON_SKELETON and pp.keypoints == cfg.COCO_KEYPOINTS
    assert pp.limb_collect.K == 32 and pp.limb_collect.resize_factor == 1.0
    assert pp.limb_group.dist_max == 40 and pp.limb_group.n_keypoints == 17
    assert pp.limb_collect.jtypes_f[:3] == [0, 0, 1] and pp.limb_collect.jtypes_t[:3] == [1, 2, 2]
    pp16 = decoder.decoder_factory(_args(headnets=['hmps', 'omp16']))
    assert pp16.skeleton == cfg.KINEMATIC_TREE_SKELETON
    with pytest.raises(Exception):
        decoder.decoder_factory(_args(headnets=['hmp', 'omp99']))
    with pytest.raises(Exception):
        decoder.decoder_factory(_args(headnets=['foo', 'omp']))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    header = open(os.path.join(ROOT, 'include', 'og_decoder.h')).read()
    declared = set(re.findall(r'\b(og_[a-z0-9_]+)\s*\(', header))
    declared -= {'og_status'}
    assert len(declared) >= 18
    lib = ctypes.CDLL(path)
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in og_decoder.h but not exported'
    assert declared == set(_lib.SIGNATURES), 'ctypes stub and header disagree'
    loaded = _lib.load()
    assert loaded.og_abi_version() == 2
    assert loaded.og_status_string(0) == b'ok'


def test_config_struct_layout_matches_header():
    # field order / types of og_config as the header declares them
    header = open(os.path.join(ROOT, 'include', 'og_decoder.h')).read()
    body = header[header.index('typedef struct og_config {'):header.index('} og_config;')]
    names = re.findall(r'^\s*(?:const\s+)?\w+\s+\*?([a-z_]+);', body, flags=re.M)
    assert names == [f[0] for f in _lib.OgConfig._fields_]


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from offsetguided_b200 import decoder
    with pytest.raises(Exception):
        decoder.hmp_NMS(torch.zeros(1, 1, 8, 8))
    with pytest.raises(Exception):
        decoder.GreedyGroup(0.06).group_skeletons(np.zeros((19, 4, 13), np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'offsetguided_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('no oracle', ''), f'{f} mentions the oracle'


def test_image_strided_views_are_passed_through():
    """Channel slices of one packed [N, 55, h, w] head output keep their memory (decoded in
    place by og_decode_features_dev_ex); other layouts are made contiguous."""
    import torch
    from offsetguided_b200.engine import _image_strided
    packed = torch.zeros(4, 55, 6, 8)
    h_view, o_view = packed[:, :17], packed[:, 17:]
    t, stride = _image_strided(h_view)
    assert t.data_ptr() == h_view.data_ptr() and stride == 55 * 48
    t, stride = _image_strided(o_view)
    assert t.data_ptr() == o_view.data_ptr() and stride == 55 * 48
    t, stride = _image_strided(packed)
    assert t.data_ptr() == packed.data_ptr() and stride == 55 * 48
    one = packed[:1, :17]
    assert _image_strided(one)[1] == 17 * 48
    transposed = packed.permute(0, 1, 3, 2)[:, :17]              # W-major planes: not passable
    t, stride = _image_strided(transposed)
    assert t.is_contiguous() and stride == 17 * 48 and t.data_ptr() != packed.data_ptr()
    every_other = packed[:, ::2]                                  # plane stride != h * w
    t, stride = _image_strided(every_other)
    assert t.is_contiguous() and stride == every_other.shape[1] * 48


def test_soft_nms_matches_reference_fixture():
    """decoder.soft_nms against outputs of the reference's own soft_nms (tests/golden/soft_nms.npz,
    written by make_golden.make_soft_nms)."""
    from offsetguided_b200 import decoder
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'soft_nms.npz'))
    at = 0
    changed = 0
    for case, n in enumerate(d['counts']):
        sub = d['inputs'][at:at + n].copy()
        ref = d['outputs'][at:at + n]
        got = decoder.soft_nms(sub, suppressed_v=0 if case % 2 == 0 else -3)
        assert got is sub and np.array_equal(got, ref), f'case {case}'
        changed += int((d['inputs'][at:at + n][..., 2] != ref[..., 2]).sum())
        at += n
    assert changed > 20
    assert decoder.soft_nms([]) == []


def test_greedy_group_pickles_without_its_handles():
    """demo_batch.py:284 passes a bound method of GreedyGroup to a process pool: the object must
    pickle; CUDA handles stay behind and are re-created by whoever uses the copy."""
    import pickle
    from offsetguided_b200 import decoder
    g = decoder.GreedyGroup(0.06, dist_max=40, use_scale=True)
    g._engines[(0, 32)] = object()                       # stands in for a live handle
    method = pickle.loads(pickle.dumps(g.group_skeletons))
    clone = method.__self__
    assert clone._engines == {} and clone.dist_max == 40 and clone.use_scale is True
    assert clone.skeleton == g.skeleton and len(g._engines) == 1


def test_abi_argument_checks_need_no_gpu():
    """Argument validation of the C ABI happens before any CUDA call: status codes, error text and
    the NULL-handle conventions can be checked on a machine without a GPU."""
    lib = _lib.load()
    assert lib.og_status_string(1) == b'invalid argument' and lib.og_status_string(5) == b'unsupported configuration'
    assert lib.og_status_string(99) == b'unknown status'
    handle = ctypes.c_void_p()
    assert lib.og_create(None, ctypes.byref(handle)) == 1 and b'null' in lib.og_last_error()
    frm, to = _lib.int32_array([0, 1]), _lib.int32_array([1, 2])

    def cfg_with(**over):
        base = dict(n_keypoints=3, n_limbs=2, limb_from=ctypes.cast(frm, _lib.c_int32_p),
                    limb_to=ctypes.cast(to, _lib.c_int32_p), topk=8, thre_hmp=0.05, min_len=0.5,
                    resize_factor=1.0, dist_max=20.0, use_scale=1, person_thre=0.05, sort_dim=2,
                    device=-1, max_images=0)
        base.update(over)
        return _lib.OgConfig(**base)
    for bad, word in ((dict(n_keypoints=0), b'n_keypoints'), (dict(n_keypoints=65), b'n_keypoints'),
                      (dict(n_limbs=0), b'n_limbs'), (dict(topk=0), b'topk'), (dict(topk=129), b'topk'),
                      (dict(sort_dim=6), b'sort_dim'),
                      (dict(limb_from=ctypes.cast(None, _lib.c_int32_p)), b'skeleton')):
        c = cfg_with(**bad)
        assert lib.og_create(ctypes.byref(c), ctypes.byref(handle)) == 1, bad
        assert word in lib.og_last_error(), (bad, lib.og_last_error())
        assert not handle.value
    far = _lib.int32_array([0, 7])                       # limb endpoint outside the keypoint list
    c = cfg_with(limb_to=ctypes.cast(far, _lib.c_int32_p))
    assert lib.og_create(ctypes.byref(c), ctypes.byref(handle)) == 1 and b'keypoint' in lib.og_last_error()
    assert lib.og_launch_count(None) == 0 and lib.og_pending(None) == 0
    assert lib.og_fused_redo_count(None) == 0 and lib.og_zero_copy_count(None) == 0
    assert lib.og_destroy(None) == 0
    assert lib.og_set_fused(None, 1) == 1 and lib.og_set_zero_copy(None, 1) == 1
