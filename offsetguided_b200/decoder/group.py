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
    """Occupancy-based suppression of duplicate keypoints (reference decoder/group.py:249-283).
    Host-side utility: the reference never calls it (the call at group.py:183 is commented
    out), it is kept only so that ``from decoder import soft_nms`` keeps working."""
    if not len(subset):
        return subset
    n_kp = len(subset[0])
    height = int(max(np.max(ann[:, 1]) for ann in subset) + 1)
    width = int(max(np.max(ann[:, 0]) for ann in subset) + 1)
    occupied = np.zeros((n_kp, height, width), dtype=np.uint8)
    for ann in subset:
        widths = np.maximum(10.0, ann[:, 3])
        assert len(occupied) == len(ann)
        for xyv, occ, jw in zip(ann[:, :3], occupied, widths):
            if xyv[2] == -1:
                continue
            x = np.clip(xyv[0], 0.0, occ.shape[1] - 1).astype(int)
            y = np.clip(xyv[1], 0.0, occ.shape[0] - 1).astype(int)
            if occ[y, x]:
                xyv[2] = suppressed_v
            else:
                scalar_square_add_single(occ, xyv[0], xyv[1], jw, 1)
    return subset


def scalar_square_add_single(field, x, y, width, value):
    """Add ``value`` to the square of half-width ``width`` round (x, y), clipped to the field and
    at least one pixel large (reference decoder/group.py:278-283; used by soft_nms)."""
    x0 = max(0, int(x - width))
    y0 = max(0, int(y - width))
    x1 = max(x0 + 1, min(field.shape[1], int(x + width) + 1))
    y1 = max(y0 + 1, min(field.shape[0], int(y + width) + 1))
    field[y0:y1, x0:x1] += value
