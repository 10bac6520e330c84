"""Offset-map helpers (reference decoder/offset.py)."""
import torch

from .. import _lib
from ..engine import DecoderEngine, as_cuda_f32, _ptr, _stream_ptr

_ENGINES = {}


def scored_offset(hmp, off, jtypes_f, jtypes_t, kernel_size=7):
    """Refine offsets with the heat-map responses round the start joint
    (reference decoder/offset.py:8-43):
    ``sumpool_k(hmp[jf] * off) / (sumpool_k(hmp[jf]) + 1e-6)`` with zero padding.

    Unlike the reference (whose ``squeeze()`` at :31 drops a batch of one) any batch
    size is accepted."""
    lib = _lib.load()
    src = hmp.device
    h = as_cuda_f32(hmp)
    o = as_cuda_f32(off, h.device)
    n, c, hh, ww = h.shape
    skeleton = tuple((int(a), int(b)) for a, b in zip(jtypes_f, jtypes_t))
    assert o.shape[1] == 2 * len(skeleton), 'offset channels must be 2 * number of limbs'
    key = (h.device.index, c, skeleton)
    if key not in _ENGINES:
        _ENGINES[key] = DecoderEngine(c, skeleton, topk=1, device=h.device)
    eng = _ENGINES[key]
    out = torch.empty_like(o)
    with torch.cuda.device(h.device):
        _lib.check(lib.og_scored_offset_f32(eng._h, _ptr(h), _ptr(o), n, hh, ww, int(kernel_size),
                                            _ptr(out), _stream_ptr(h.device)))
    return out.to(src)


def pack_jtypes(skeleton):
    """(from-joint list, to-joint list) of a skeleton (reference decoder/offset.py:46-51)."""
    return [a for a, _ in skeleton], [b for _, b in skeleton]
