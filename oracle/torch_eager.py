"""TEST / BENCH INFRASTRUCTURE — the decoding path written with stock PyTorch operators.

A second baseline for bench.py ("the same reference code on CUDA tensors on one
B200"): what the path costs when every stage is an eager ATen kernel on the GPU, the way the
reference runs it (decoder/factory.py:52-96), followed by the reference's CPU grouping (here the
multi-threaded C oracle, which is kinder to this baseline than the reference's process pool).
The reference sources cannot travel to the GPU box, so this is an independent formulation of
the same stages from SURVEY.md 8a — flip fusion, F.interpolate, max-pool NMS, torch.topk, offset
gather, nearest to-candidate, limb score.  Never imported by the product package.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _fuse_flipped(hmp, omp, kp_flips, limb_flips, limb_reserve):
    """Average the originals with their mirrored copies (heat maps: swap left / right channels;
    offsets: mirror, negate x, permute limbs, keep self-mirrored limbs un-averaged)."""
    n = hmp.shape[0] // 2
    heat = (hmp[:n] + hmp[n:].flip(-1)[:, kp_flips]) * 0.5
    pairs = omp.reshape(2 * n, -1, 2, omp.shape[-2], omp.shape[-1])
    mirrored = pairs[n:].flip(-1).clone()
    mirrored[:, :, 0].neg_()
    offs = (pairs[:n] + mirrored[:, limb_flips]) * 0.5
    offs[:, limb_reserve] = pairs[:n][:, limb_reserve]
    return heat, offs.reshape(n, -1, omp.shape[-2], omp.shape[-1])


def peak_candidates(heat, k):
    """3 x 3 max-pool NMS with zero padding, then the k best responses of every channel."""
    padded = F.pad(heat, (1, 1, 1, 1), value=0.0)
    keep = F.max_pool2d(padded, 3, stride=1) == heat
    scores, index = torch.topk((heat * keep).flatten(2), k)
    return scores, index


def limb_table(heat, offs, skeleton, k, thre_hmp, min_len, ratio=1.0):
    """(N, L, K, 13) limb candidates of full-resolution maps."""
    n, c, h, w = heat.shape
    scores, index = peak_candidates(heat, k)
    jf = torch.tensor([a for a, _ in skeleton], device=heat.device)
    jt = torch.tensor([b for _, b in skeleton], device=heat.device)
    xy = torch.stack((index % w, index // w), dim=-1)
    xy = torch.where((scores < thre_hmp).unsqueeze(-1), xy - 100000, xy).float()
    s_f, s_t, i_f, i_t, p_f, p_t = scores[:, jf], scores[:, jt], index[:, jf], index[:, jt], xy[:, jf], xy[:, jt]
    pairs = offs.reshape(n, len(skeleton), 2, h * w)
    vec = torch.gather(pairs, 3, i_f.unsqueeze(2).expand(-1, -1, 2, -1)).transpose(2, 3)
    guided = p_f + vec * ratio
    dist = (guided.unsqueeze(3) - p_t.unsqueeze(2)).norm(dim=-1)
    best, arg = dist.min(dim=-1)
    m_s = torch.gather(s_t, 2, arg)
    m_i = torch.gather(i_t, 2, arg)
    m_p = torch.gather(p_t, 2, arg.unsqueeze(-1).expand(-1, -1, -1, 2))
    length = (p_f - m_p).norm(dim=-1).clamp(min=min_len)
    limb_score = s_f * m_s * torch.exp(-best / length)
    ids_f = (i_f + (jf * h * w).view(1, -1, 1)).float()
    ids_t = (m_i + (jt * h * w).view(1, -1, 1)).float()
    four = torch.full_like(s_f, 4.0)
    return torch.stack((p_f[..., 0], p_f[..., 1], s_f, m_p[..., 0], m_p[..., 1], m_s, ids_f, ids_t, best,
                        length, limb_score, four, four), dim=-1)


def generate_poses(hmp, omp, skeleton, n_keypoints, *, topk, thre_hmp, min_len, person_thre, dist_max,
                   use_scale=True, stride=4, resize_mode='bicubic', flip_test=False, kp_flips=None,
                   limb_flips=None, limb_reserve=None, group=None):
    """Eager-PyTorch decode of a batch of network-resolution maps (tensors on any device).
    ``group`` maps a (N, L, K, 13) numpy array to the list of pose arrays (default: C oracle)."""
    if flip_test:
        hmp, omp = _fuse_flipped(hmp, omp, list(kp_flips), list(limb_flips), list(limb_reserve))
    if stride > 1:
        hmp = F.interpolate(hmp, scale_factor=stride, mode=resize_mode)
        omp = F.interpolate(omp, scale_factor=stride, mode='bilinear')
    limbs = limb_table(hmp, omp, skeleton, topk, thre_hmp, min_len).cpu().numpy()
    if group is None:
        from oracle import c_oracle
        return c_oracle.group_batch(limbs, skeleton, n_keypoints, person_thre, 2, dist_max, use_scale)
    return group(np.ascontiguousarray(limbs))
