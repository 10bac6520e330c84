"""Image sharding across GPUs.

Every decoder stage is per image, so a batch is split into contiguous chunks, one per rank
(one process per GPU); no collective is needed to decode.  With flip-test inputs — N originals
followed by their N mirrored copies (reference evaluate.py:211-212) — image i and its copy i + N
stay on the same rank.  ``gather_poses`` is an optional convenience for callers that want every
result on one rank; it moves small pose arrays only and is not part of the decoding path.
"""
import torch
import torch.distributed as dist


def partition(n_images, world_size):
    """Contiguous [start, stop) chunk of every rank, ceil(N / G) images each."""
    per = (n_images + world_size - 1) // world_size if world_size > 0 else n_images
    return [(min(r * per, n_images), min((r + 1) * per, n_images)) for r in range(world_size)]


def shard_batch(t, rank, world_size, flip_test=False):
    """This rank's slice of a batched tensor (N, ...) or, with flip_test, (2N, ...)."""
    n = t.shape[0] // 2 if flip_test else t.shape[0]
    start, stop = partition(n, world_size)[rank]
    if not flip_test:
        return t[start:stop]
    return torch.cat((t[start:stop], t[n + start:n + stop]), dim=0)


def shard_features(features, rank, world_size, flip_test=False, hmp_index=0, omp_index=1, feat_stage=-1):
    """Slice the nested ``features`` structure PostProcess.generate_poses takes
    (reference decoder/factory.py:54-63); non-tensor heads are passed through."""
    def cut(x):
        return shard_batch(x, rank, world_size, flip_test) if isinstance(x, torch.Tensor) else x
    out = []
    for head in features:
        out.append([[cut(x) for x in stages] if isinstance(stages, (list, tuple)) else stages
                    for stages in head])
    return out


def gather_poses(local_poses, n_images, group=None, dst=0):
    """Collect every rank's pose list on ``dst`` in image order (None elsewhere)."""
    if not dist.is_available() or not dist.is_initialized():
        return list(local_poses)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(list(local_poses), bucket, dst=dst, group=group)
    if rank != dst:
        return None
    merged = []
    for (start, stop), part in zip(partition(n_images, world), bucket):
        assert len(part) == stop - start, 'a rank returned the wrong number of images'
        merged.extend(part)
    return merged


def decode_sharded(post, features, flip_test=False, group=None, gather=True):
    """Decode this rank's images with ``post`` (a PostProcess) and optionally gather."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    hmps = features[post.hmp_index][0][post.feat_stage]
    n = hmps.shape[0] // 2 if flip_test else hmps.shape[0]
    local = post.generate_poses(shard_features(features, rank, world, flip_test), flip_test=flip_test)
    return gather_poses(local, n, group) if gather else local
