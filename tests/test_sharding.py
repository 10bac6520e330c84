"""CPU suite: image sharding logic with world-size-2 gloo process groups."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from offsetguided_b200 import sharding


def test_partition_covers_everything():
    for n in (0, 1, 7, 8, 64, 65):
        for g in (1, 2, 4, 8):
            parts = sharding.partition(n, g)
            assert len(parts) == g
            covered = [i for s, e in parts for i in range(s, e)]
            assert covered == list(range(n))
            assert max(e - s for s, e in parts) <= (n + g - 1) // g


def test_flip_pairs_stay_together():
    n = 6
    t = torch.arange(2 * n).view(2 * n, 1)
    seen = []
    for r in range(4):
        part = sharding.shard_batch(t, r, 4, flip_test=True)
        k = part.shape[0] // 2
        for i in range(k):
            assert int(part[k + i]) == int(part[i]) + n
        seen += [int(v) for v in part[:k, 0]]
    assert seen == list(range(n))


class _StubPost(object):
    """Stands in for PostProcess on the CPU box: 'decodes' an image into an array that
    encodes which image it was (the GPU decode itself is covered by tests -m gpu)."""
    hmp_index, omp_index, feat_stage = 0, 1, -1

    def generate_poses(self, features, flip_test=False):
        hmps = features[0][0][-1]
        n = hmps.shape[0] // 2 if flip_test else hmps.shape[0]
        out = []
        for i in range(n):
            tag = float(hmps[i].flatten()[0])
            mirror = float(hmps[n + i].flatten()[0]) if flip_test else -1.0
            out.append(np.full((int(tag) % 3, 17, 6), tag + mirror, np.float32))
        return out


def _worker(rank, world, port, flip, n, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        total = 2 * n if flip else n
        hmp = torch.arange(total, dtype=torch.float32).view(total, 1, 1, 1).repeat(1, 17, 2, 2)
        off = torch.zeros(total, 38, 2, 2)
        feats = [[[hmp], [[]], [[]]], [[off], [[]], [[]]]]
        merged = sharding.decode_sharded(_StubPost(), feats, flip_test=flip)
        if rank == 0:
            assert len(merged) == n
            for i, p in enumerate(merged):
                expect = float(i) + (float(n + i) if flip else -1.0)
                assert p.shape == (i % 3, 17, 6)
                assert np.all(p == np.float32(expect))
            ret.put('ok')
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('flip,n', [(False, 5), (True, 4)])
def test_decode_sharded_world2_gloo(flip, n):
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + (7 if flip else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, flip, n, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 'ok'
