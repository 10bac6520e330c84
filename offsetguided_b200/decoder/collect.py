"""Generate and collect all candidate keypoint pairs (reference decoder/collect.py)."""
import logging

import torch

from ..config import COCO_KEYPOINTS, COCO_PERSON_SKELETON
from ..engine import DecoderEngine, as_cuda_f32

LOG = logging.getLogger(__name__)


class LimbsCollect(object):
    """Collect all candidate keypoints and pair them into limbs on the basis of the
    guiding offset vectors (reference decoder/collect.py:15-60, same constructor).

    ``generate_limbs`` runs two CUDA kernels: K1 (fused NMS + threshold + top-K) and
    K2 (offset gather, K x K nearest-candidate search, limb score).

    Attributes mirror the reference: hmp_s, off_s, resize_factor, keypoints, skeleton,
    K, thre_hmp, min_len, include_jitter_offset, include_scale, use_jitter_offset,
    jtypes_f, jtypes_t.
    """

    def __init__(self, hmp_s, off_s, *, topk=40, thre_hmp=0.08, min_len=3,
                 include_jitter_offset=False, include_scale=False, use_jitter_offset=True,
                 keypoints=COCO_KEYPOINTS, skeleton=COCO_PERSON_SKELETON):
        LOG.info('number of skeleton limbs: %d, response threshold to drop keypoints: '
                 'threshold=%.4f', len(skeleton), thre_hmp)
        self.hmp_s = hmp_s
        self.off_s = off_s
        self.resize_factor = off_s / hmp_s
        self.keypoints = keypoints
        self.skeleton = skeleton
        self.K = topk
        self.thre_hmp = thre_hmp
        self.min_len = min_len
        self.include_jitter_offset = include_jitter_offset
        self.include_scale = include_scale
        self.use_jitter_offset = use_jitter_offset
        self.jtypes_f, self.jtypes_t = self.pack_jtypes(skeleton)
        self._engines = {}

    def _engine(self, device):
        key = (device.type, device.index)
        if key not in self._engines:
            self._engines[key] = DecoderEngine(
                len(self.keypoints), self.skeleton, topk=self.K, thre_hmp=self.thre_hmp,
                min_len=self.min_len, resize_factor=self.resize_factor, device=device)
        return self._engines[key]

    def generate_limbs(self, hmps_hr, jomps_hr, offs_hr, scmps_hr, vector_nd=2):
        """All candidate limbs of a batch (reference decoder/collect.py:62-236).

        Args:
            hmps_hr (Tensor): (N, C, H, W) heat maps at decode resolution.
            jomps_hr: (N, 2, H, W) jitter-offset maps, used when ``include_jitter_offset``.
            offs_hr (Tensor): (N, 2L, H, W) guiding offsets, (x, y) interleaved per limb
                ((N, 4L, H, W) with ``vector_nd=4``).
            scmps_hr: (N, C, H, W) keypoint-scale maps, used when ``include_scale``.
            vector_nd (int): 2, or 4 for the flip-concatenated offsets of ``cat_flip_offs``.

        Returns:
            Tensor (N, L, K, 13) float32 on the device of ``hmps_hr``:
            [x1, y1, v1, x2, y2, v2, ind1, ind2, min_dist, len, limb_score, scale1, scale2].
            Rows whose from-candidate scores below ``thre_hmp`` are inert placeholders
            (moved 100000 px off the image exactly as in the reference).
        """
        assert hmps_hr.shape[-2:] == offs_hr.shape[-2:], 'spatial resolution should be equal'
        if vector_nd not in (2, 4):
            raise ValueError('vector_nd must be 2, or 4 for flip-concatenated offsets')
        src = hmps_hr.device
        heat = as_cuda_f32(hmps_hr)
        eng = self._engine(heat.device)
        scores, inds, _ = eng.nms_topk(heat)
        scales = scmps_hr if (self.include_scale and isinstance(scmps_hr, torch.Tensor)) else None
        jomps = jomps_hr if (self.include_jitter_offset and isinstance(jomps_hr, torch.Tensor)) else None
        limbs = eng.limb_score(scores, inds, offs_hr, scales, jomps, vector_nd,
                               self.use_jitter_offset)
        return limbs.to(src)

    @staticmethod
    def pack_jtypes(skeleton):
        """(from-joint list, to-joint list) (reference decoder/collect.py:238-244)."""
        return [a for a, _ in skeleton], [b for _, b in skeleton]
