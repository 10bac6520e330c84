"""BASELINE configs[4]: a bf16 network forward feeding the decoder, batch 256 over 8 GPUs
(32 images + their mirrored copies per GPU), decode of batch i overlapping the forward of batch i + 1.

The reference's Hourglass4Stage is not part of this repository (and must not be vendored); the
producer here is a stand-in conv stack with the same hand-over: ONE packed bf16 tensor
[2N, 55, 160, 160] (models/hourglass_4stage.py:58,119-121: oup_dim = 17 heat + 38 offset channels
at stride 4), whose channel slices the decoder reads in place (og_decode_features_dev_ex).  A
random-init network emits ~0 heat maps (no candidates), so pre-rendered synthetic maps are added to
the head output: the decoder sees 6 persons per image.

    python profiles/tools/config5_pipeline.py                       # one GPU, 32 images per batch
    torchrun --nproc-per-node 8 profiles/tools/config5_pipeline.py  # batch 256 over 8 GPUs
Reports forward-only, decode-only and pipelined time per batch: the decode-added time per batch is
pipelined - forward-only.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch                    # noqa: E402
import torch.distributed as dist        # noqa: E402
import bench                    # noqa: E402


class Block(torch.nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.a = torch.nn.Conv2d(ch, ch, 3, padding=1, bias=False)
        self.b = torch.nn.Conv2d(ch, ch, 3, padding=1, bias=False)
        self.n1 = torch.nn.BatchNorm2d(ch)
        self.n2 = torch.nn.BatchNorm2d(ch)

    def forward(self, x):
        y = torch.relu(self.n1(self.a(x)))
        return torch.relu(x + self.n2(self.b(y)))


class StandIn(torch.nn.Module):
    """stride-4 stem, four residual blocks at 160 x 160 x 128, 1 x 1 head with 55 channels"""
    def __init__(self, ch=128, out=55):
        super().__init__()
        self.stem = torch.nn.Sequential(torch.nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False), torch.nn.ReLU(),
                                        torch.nn.Conv2d(64, ch, 3, stride=2, padding=1, bias=False), torch.nn.ReLU())
        self.blocks = torch.nn.Sequential(*[Block(ch) for _ in range(4)])
        self.head = torch.nn.Conv2d(ch, out, 1)

    def forward(self, x):
        return self.head(self.blocks(self.stem(x)))


class A(object):
    workload, batch, long_edge, no_flip = 'cfg2', 0, 0, False


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
            os.environ.pop('NCCL_DEBUG')
        dist.init_process_group('nccl', device_id=dev)
    n = int(os.environ.get('CFG5_IMAGES_PER_GPU', '32'))
    steps = int(os.environ.get('CFG5_STEPS', '30'))
    w = bench.workload(A())
    post = bench.make_post(w, n)
    torch.backends.cudnn.benchmark = True
    net = StandIn()
    with torch.no_grad():                  # like the reference's normal(0, 0.001) init: outputs ~ 0
        net.head.weight.mul_(1e-3)
        net.head.bias.zero_()
    net = net.to(dev).to(torch.bfloat16).eval().to(memory_format=torch.channels_last)
    hmp, omp = bench.lowres_inputs(5000 + rank, n, 640, True, w)
    synth = torch.cat((torch.from_numpy(hmp), torch.from_numpy(omp)), dim=1).to(dev).to(torch.bfloat16)
    images = torch.randn(2 * n, 3, 640, 640, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    outs = [torch.empty(2 * n, 55, 160, 160, device=dev, dtype=torch.bfloat16) for _ in range(2)]

    def forward(i):
        with torch.no_grad():
            y = net(images)
            outs[i % 2].copy_(y)            # the packed NCHW hand-over buffer of batch i
            outs[i % 2].add_(synth)
        return outs[i % 2]

    def feats(o):
        return [[[o[:, :17]], [[]], [[]]], [[o[:, 17:]], [[]], [[]]]]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(4):
        poses = post.generate_poses(feats(forward(i)), flip_test=True)
    persons = sum(len(p) for p in poses)
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        forward(i)
    torch.cuda.synchronize(dev)
    t_fwd = (time.perf_counter() - t0) / steps
    barrier()
    o = forward(0)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(steps):
        post.generate_poses(feats(o), flip_test=True)
    t_dec = (time.perf_counter() - t0) / steps
    barrier()
    t0 = time.perf_counter()
    post.submit(feats(forward(0)), flip_test=True)
    for i in range(1, steps):
        post.submit(feats(forward(i)), flip_test=True)      # forward i is queued; decode i waits for it on the GPU
        post.collect()                                      # poses of batch i - 1 while forward i runs
    post.collect()
    torch.cuda.synchronize(dev)
    t_pipe = (time.perf_counter() - t0) / steps
    barrier()
    t = torch.tensor([t_fwd, t_dec, t_pipe], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_fwd, t_dec, t_pipe = [float(v) for v in t.cpu()]
    if rank == 0:
        print(json.dumps({
            'config': 'BASELINE configs[4]: bf16 forward (stand-in conv stack, packed [2N, 55, 160, 160] hand-over) + decoder, '
                      '%d images per GPU x %d GPUs = batch %d, flip-test' % (n, world, n * world),
            'forward_only_ms_per_batch': round(1e3 * t_fwd, 4),
            'decode_only_ms_per_batch (synchronous generate_poses)': round(1e3 * t_dec, 4),
            'pipelined_ms_per_batch (submit / collect, decode i beside forward i + 1)': round(1e3 * t_pipe, 4),
            'decode_added_ms_per_batch': round(1e3 * (t_pipe - t_fwd), 4),
            'images_per_s_pipeline': round(n * world / t_pipe), 'images_per_s_forward_only': round(n * world / t_fwd),
            'persons_per_batch_per_gpu': persons}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
