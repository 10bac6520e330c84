"""Greedily group keypoints into persons (reference decoder/group.py)."""
import logging

import numpy as np
import torch

from ..config import COCO_KEYPOINTS, COCO_PERSON_SKELETON
from ..engine import DecoderEngine

LOG = logging.getLogger(__name__)


class GreedyGroup(object):
    """Greedily group the limbs of ONE image into individual skeletons
    (reference decoder/group.py:16-37, same constructor).  ``group_skeletons`` runs the
    one-CTA-per-image CUDA kernel K3; ``group_batch`` does a whole batch in one launch.

    Args:
        person_thre (float): threshold on the pose-instance score.
        sort_dim (int): pose column the instance score is computed from
            (2 = keypoint score, 4 = limb score).
        dist_max (float): limbs whose guided end misses by more than this are dropped.
        use_scale (bool): gate with max(dist_max, keypoint scale) instead.
    """

    def __init__(self, person_thre, *, sort_dim=2, dist_max=10, use_scale=False,
                 keypoints=COCO_KEYPOINTS, skeleton=COCO_PERSON_SKELETON):
        self.person_thre = person_thre
        self.use_scale = use_scale
        self.sort_dim = sort_dim
        self.skeleton = skeleton
        self.keypoints = keypoints
        self.dist_max = dist_max
        self.n_keypoints = len(keypoints)
        self._engines = {}

    def __getstate__(self):
        # The reference hands bound methods of this class to a multiprocessing pool
        # (demo_batch.py:284); handles of the CUDA library do not travel, a copy re-creates its own.
        state = dict(self.__dict__)
        state['_engines'] = {}
        return state

    def _engine(self, topk, device=None):
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        key = (device.index, int(topk))
        if key not in self._engines:
            self._engines[key] = DecoderEngine(
                self.n_keypoints, self.skeleton, topk=topk, dist_max=self.dist_max,
                use_scale=self.use_scale, person_thre=self.person_thre, sort_dim=self.sort_dim,
                device=device)
        return self._engines[key]

    def group_skeletons(self, limbs):
        """Group the candidate limbs of a single image (reference decoder/group.py:39-185).

        Args:
            limbs (np.ndarray | Tensor): (L, K, 13).
        Returns:
            np.ndarray (M, C, 6) float32: [x, y, v, s, limb_score, ind] per keypoint.
        """
        assert len(limbs) == len(self.skeleton), 'check the skeleton config and input limbs Tensor'
        return self.group_batch(limbs[None])[0]

    def group_batch(self, limbs):
        """(N, L, K, 13) -> list of N pose arrays, one kernel launch."""
        if isinstance(limbs, torch.Tensor):
            t = limbs
            device = t.device if t.is_cuda else None
        else:
            t = torch.from_numpy(np.ascontiguousarray(limbs, dtype=np.float32))
            device = None
        assert t.shape[1] == len(self.skeleton), 'check the skeleton config and input limbs Tensor'
        return self._engine(t.shape[2], device).group(t)


def soft_nms(subset, suppressed_v=0):
    """Suppress duplicate keypoints (reference decoder/group.py:249-283; the reference never calls
    it — the call at group.py:183 is commented out — it is kept so that ``from decoder import
    soft_nms`` keeps working).  Poses are visited in order; a keypoint whose (clipped, truncated)
    position falls into the square of half-width max(10, scale) round an EARLIER surviving
    keypoint of the same type gets ``v = suppressed_v``.  ``subset`` is modified in place.

    Instead of rasterising the squares into a (C, H, W) occupancy field, the accepted squares of a
    keypoint type are kept as a list and tested directly (the same integer bounds, clipped to the
    extent the reference's field would have)."""
    if not len(subset):
        return subset
    n_kp = len(subset[0])
    field_h = int(max(np.max(pose[:, 1]) for pose in subset) + 1)
    field_w = int(max(np.max(pose[:, 0]) for pose in subset) + 1)
    for joint in range(n_kp):
        lo_x, hi_x, lo_y, hi_y = [], [], [], []           # half-open pixel ranges of the accepted squares
        for pose in subset:
            assert len(pose) == n_kp
            px, py, v = pose[joint, 0], pose[joint, 1], pose[joint, 2]
            if v == -1:
                continue
            half = max(10.0, float(pose[joint, 3]))
            cx = int(np.clip(px, 0.0, field_w - 1))
            cy = int(np.clip(py, 0.0, field_h - 1))
            if any(a <= cx < b and c <= cy < d for a, b, c, d in zip(lo_x, hi_x, lo_y, hi_y)):
                pose[joint, 2] = suppressed_v
                continue
            x0, y0 = max(0, int(px - half)), max(0, int(py - half))
            lo_x.append(x0)
            lo_y.append(y0)
            hi_x.append(min(field_w, max(x0 + 1, min(field_w, int(px + half) + 1))))
            hi_y.append(min(field_h, max(y0 + 1, min(field_h, int(py + half) + 1))))
    return subset
