"""Heat-map peak extraction (reference decoder/heatmap.py)."""
import ctypes
import logging

import torch

from .. import _lib
from ..engine import DecoderEngine, as_cuda_f32, _ptr, _stream_ptr

LOG = logging.getLogger(__name__)

_UTILITY = {}


def _utility_engine(device):
    """A minimal handle for the stand-alone helpers (they need no skeleton)."""
    key = (device.type, device.index)
    if key not in _UTILITY:
        _UTILITY[key] = DecoderEngine(1, [(0, 0)], topk=1, device=device)
    return _UTILITY[key]


def _on_cuda(t):
    src = t.device
    out = as_cuda_f32(t)
    return out, src


def normalize_hmps():
    """Placeholder of the reference (decoder/heatmap.py:10-12: "todo: filter and smooth the
    heatmaps", never implemented there); kept so that the module exports the same names."""
    return None


def hmp_NMS(heat, kernel=3):
    """3x3 max-pool NMS (reference decoder/heatmap.py:15-35): peaks keep their value,
    every other response becomes 0.  The border is zero padded, plateaus survive.

    Args:
        heat (Tensor): (N, C, H, W) float32.
        kernel: only 3 is implemented (the reference never uses another size).
    """
    if kernel != 3:
        raise NotImplementedError('hmp_NMS: only the 3x3 window of the reference is implemented')
    lib = _lib.load()
    x, src = _on_cuda(heat)
    n, c, h, w = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(lib.og_hmp_nms_f32(_ptr(x), _ptr(out), n, c, h, w, _stream_ptr(x.device)))
    return out.to(src)


def topK_channel(scores, K=40):
    """Per-channel top-K over H*W (reference decoder/heatmap.py:38-49).

    Returns (topk_scores f32, topk_idxs i64, topk_ys i64, topk_xs i64), each (N, C, K),
    ordered by (score desc, flat index asc).  ``ys = idx // w`` as under the reference's
    pinned torch 1.3.1.  Top K may include very small, even zero, responses."""
    x, src = _on_cuda(scores)
    n, c, h, w = x.shape
    eng = _utility_engine(x.device)
    s, i = eng.topk_channel(x, K)
    idx = i.to(torch.int64)
    return s.to(src), idx.to(src), (idx // w).to(src), (idx % w).to(src)


def joint_dets(hmps, k):
    """Top-k candidate keypoints of every heat-map channel
    (reference decoder/heatmap.py:52-59): hmp_NMS followed by topK_channel."""
    return topK_channel(hmp_NMS(hmps), K=k)
