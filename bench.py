#!/usr/bin/env python
"""Benchmark of the OffsetGuided post-network decoder on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" decodes one batch of synthetic network outputs: BASELINE.json configs[1]
settings (COCO skeleton, flip-test fusion, x4 resize to 640x640, topk 32, thre-hmp 0.04,
person-thre 0.04, dist-max 40) at the north-star batch of 64 images per GPU.

  value     images/s of the hot path (K1 NMS+top-K -> K2 limb scoring -> K3 grouping, poses
            copied to pinned memory and fetched) on full-resolution maps RESIDENT IN HBM, two
            batches in flight, CUDA-event timed on the launching stream, max over ranks;
  e2e       images/s through the reference-facing API PostProcess.generate_poses with
            HOST (pinned) network-resolution maps: H2D copy, flip fusion, x4 resize,
            K1..K3 and the D2H read of the poses are all inside the timed region;
  features_dev  the same decode from DEVICE-resident network-resolution maps (the call
            evaluate.py makes right after model(images)), extra key;
  roofline  K1 (both passes) algorithmic bytes N*C*H*W*4 over its CUDA-event duration,
            against the measured HBM copy peak (MEASURED_PEAKS.json);
  cpu_baseline  the oracle port of the reference decoder (oracle/) on the host cores, on
            a bounded sample of the same workload.

Multi-GPU: images are independent, every rank decodes its own batch (weak scaling, no
collective on the data path); launched by torchrun for N > 1.
`--impl reference` times the CPU oracle port alone (rank 0 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'decoded images/s @640 long-edge'
UNIT = 'images/s'
TOPK, THRE_HMP, PERSON_THRE, DIST_MAX, MIN_LEN = 32, 0.04, 0.04, 40.0, 0.5
PERSONS = 6


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='images per GPU per step')
    ap.add_argument('--long-edge', type=int, default=640)
    ap.add_argument('--no-flip', action='store_true', help='e2e without flip-test inputs')
    ap.add_argument('--cpu-sample', type=int, default=0, help='images of the CPU sample (0 = auto)')
    ap.add_argument('--hot-only', action='store_true',
                    help='run only the HBM-resident hot path (for ncu captures; prints no bench line)')
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {
        'workload': 'BASELINE configs[1] settings at batch %d per GPU: COCO 17 keypoints / 19 limbs, '
                    'network maps %dx%d (x4 -> %dx%d), %s, topk=%d, thre_hmp=%.2f, person_thre=%.2f, '
                    'dist_max=%d' % (args.batch, args.long_edge // 4, args.long_edge // 4, args.long_edge,
                                     args.long_edge, 'no flip' if args.no_flip else 'flip-test fusion',
                                     TOPK, THRE_HMP, PERSON_THRE, int(DIST_MAX)),
        'images_per_gpu_per_step': args.batch,
        'global_batch': args.batch * n_gpus,
        'parallelism': 'image-sharded x%d, no collective' % n_gpus,
        'l2_policy': 'hot-path inputs are %.1f GB per step (heat %.2f GB + offsets %.2f GB) >> 126 MB L2; '
                     'e2e / features_dev read %.0f MB of network-resolution heat maps per step'
                     % (args.batch * 55 * args.long_edge ** 2 * 4 / 1e9, args.batch * 17 * args.long_edge ** 2 * 4 / 1e9,
                        args.batch * 38 * args.long_edge ** 2 * 4 / 1e9,
                        args.batch * (1 if args.no_flip else 2) * 17 * (args.long_edge // 4) ** 2 * 4 / 1e6),
        'persons_per_image': PERSONS,
    }


# --------------------------------------------------------------------------- inputs
def lowres_inputs(seed, n, long_edge, flip):
    """Network-resolution maps as the reference encoder renders them (oracle/scenes.py
    restates encoder/heatmap.py and encoder/offset.py); flipped half from mirrored persons."""
    from oracle import scenes
    from offsetguided_b200 import config as cfg
    skel = cfg.COCO_PERSON_SKELETON
    kp = cfg.heatmap_hflip(cfg.COCO_KEYPOINTS)
    distinct = min(n, 16)
    hs, os_, hf, of = [], [], [], []
    for i in range(distinct):
        rng = np.random.RandomState(seed + i)
        p = scenes.make_persons(rng, PERSONS, long_edge, long_edge,
                                scale_range=(long_edge / 64.0, long_edge / 27.0))
        hs.append(scenes.render_heatmaps(p, long_edge, long_edge) +
                  rng.uniform(0, 0.02, size=(17, long_edge // 4, long_edge // 4)).astype(np.float32))
        os_.append(scenes.render_offsets(p, long_edge, long_edge, skel))
        if flip:
            pf = scenes.mirror_persons(p, long_edge, kp)
            hf.append(scenes.render_heatmaps(pf, long_edge, long_edge) +
                      rng.uniform(0, 0.02, size=(17, long_edge // 4, long_edge // 4)).astype(np.float32))
            of.append(scenes.render_offsets(pf, long_edge, long_edge, skel))
    reps = (n + distinct - 1) // distinct

    def tile(lst):
        return np.concatenate([np.stack(lst)] * reps)[:n]
    hmp = tile(hs)
    omp = tile(os_)
    if flip:
        hmp = np.concatenate((hmp, tile(hf)))
        omp = np.concatenate((omp, tile(of)))
    omp[~np.isfinite(omp)] = 0
    return hmp.astype(np.float32), omp.astype(np.float32)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons of one GPU during the timed region."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None,
                'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# --------------------------------------------------------------------------- CPU arm
def cpu_decode_fn():
    """The oracle port of the reference decoder: multi-threaded C restatement when it has
    been built (oracle/og_oracle.c), else the numpy restatement (single thread)."""
    from offsetguided_b200 import config as cfg
    skel = cfg.COCO_PERSON_SKELETON
    kp = cfg.heatmap_hflip(cfg.COCO_KEYPOINTS)
    fl, rs = cfg.offset_hflip(cfg.COCO_KEYPOINTS, skel)
    try:
        from oracle import c_oracle
        c_oracle.load()
        cores = c_oracle.num_threads()

        def run(hmp, omp, flip, stage_seconds=None):
            return c_oracle.generate_poses(hmp, omp, skel, 17, topk=TOPK, thre_hmp=THRE_HMP,
                                           min_len=MIN_LEN, person_thre=PERSON_THRE, dist_max=DIST_MAX,
                                           use_scale=True, flip_test=flip, kp_flips=kp, limb_flips=fl,
                                           limb_reserve=rs, stage_seconds=stage_seconds)
        return run, cores, 'oracle/og_oracle.c (C + OpenMP restatement of the reference decoder)'
    except Exception:
        from oracle import ref_oracle as ro

        def run(hmp, omp, flip, stage_seconds=None):
            return ro.generate_poses(hmp, omp, skel, 17, topk=TOPK, thre_hmp=THRE_HMP, min_len=MIN_LEN,
                                     person_thre=PERSON_THRE, dist_max=DIST_MAX, use_scale=True,
                                     flip_test=flip, kp_flips=kp, limb_flips=fl, limb_reserve=rs)
        return run, 1, 'oracle/ref_oracle.py (numpy restatement of the reference decoder)'


def time_cpu(args, n_images, repeats):
    run, cores, what = cpu_decode_fn()
    flip = not args.no_flip
    hmp, omp = lowres_inputs(9000, n_images, args.long_edge, flip)
    run(hmp[:1] if not flip else np.concatenate((hmp[:1], hmp[n_images:n_images + 1])),
        omp[:1] if not flip else np.concatenate((omp[:1], omp[n_images:n_images + 1])), flip)   # warm
    times, stage_runs = [], []
    for _ in range(repeats):
        stage = {}
        t0 = time.perf_counter()
        poses = run(hmp, omp, flip, stage)
        times.append(time.perf_counter() - t0)
        stage_runs.append(stage)
        assert len(poses) == n_images
    best = min(times)
    stage_ms = {k: 1e3 * v / n_images for k, v in stage_runs[times.index(best)].items()}
    return n_images / best, cores, what, times, stage_ms


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    run, cores, what = cpu_decode_fn()
    n_images = args.cpu_sample or (8 if cores > 1 else 1)
    flip = not args.no_flip
    hmp, omp = lowres_inputs(9000, n_images, args.long_edge, flip)
    for _ in range(min(args.warmup, 1)):
        run(hmp, omp, flip)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(hmp, omp, flip)
    dt = time.perf_counter() - t0
    value = n_images * args.steps / dt
    sample = '%d images per step (of the %d-image batch), %d steps, %s' % (n_images, args.batch, args.steps, what)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': min(args.warmup, 1), 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': workload_config(args, args.gpus),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from offsetguided_b200 import config as cfg
    from offsetguided_b200 import decoder
    from offsetguided_b200.engine import DecoderEngine
    from oracle import scenes

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # This image exports NCCL_DEBUG=VERSION, which makes NCCL print its version banner on
        # STDOUT (also at WARN; probed) — stdout carries the one JSON line only.  Any other explicit
        # setting of the caller (INFO, TRACE) is left alone.
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
            os.environ.pop('NCCL_DEBUG')
        dist.init_process_group('nccl', device_id=dev)
    n_gpus = world

    skel = cfg.COCO_PERSON_SKELETON
    B, E = args.batch, args.long_edge
    flip = not args.no_flip

    # ---- hot path on HBM-resident full-resolution maps
    distinct = min(B, 8)
    heat_np, offs_np = scenes.synth_hires_batch(1000 * (rank + 1), distinct, PERSONS, E, E, skel)
    heat = torch.from_numpy(heat_np).to(dev).repeat((B + distinct - 1) // distinct, 1, 1, 1)[:B].contiguous()
    offs = torch.from_numpy(offs_np).to(dev).repeat((B + distinct - 1) // distinct, 1, 1, 1)[:B].contiguous()
    del heat_np, offs_np
    eng = DecoderEngine(17, skel, topk=TOPK, thre_hmp=THRE_HMP, min_len=MIN_LEN, dist_max=DIST_MAX,
                        use_scale=True, person_thre=PERSON_THRE, device=dev)
    eng.enable_stage_timing(True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        poses = eng.decode_maps(heat, offs)
    n_persons = sum(len(p) for p in poses)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = eng.launch_count
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stages = []
    # two batches in flight: batch i + 1 is launched before batch i is fetched, so the GPU never
    # waits for the host; every batch's poses are fetched (K launches, K fetches)
    start.record()
    eng.decode_maps(heat, offs, fetch=False)
    for _ in range(args.steps - 1):
        eng.decode_maps(heat, offs, fetch=False)
        eng.fetch(B)
        stages.append(eng.last_stage_times_ms())
    poses = eng.fetch(B)
    stages.append(eng.last_stage_times_ms())
    stop.record()
    barrier()
    hot_ms = start.elapsed_time(stop)
    hot_launches = eng.launch_count - launches0
    if args.hot_only:
        sampler.stop()
        print('hot-only: %.3f ms/step, stages %s' % (hot_ms / args.steps, stages[-1]), file=sys.stderr)
        return

    # ---- end to end through the reference-facing API, host buffers
    ap = argparse.ArgumentParser()
    decoder.decoder_cli(ap)
    dargs = ap.parse_args(['--topk', str(TOPK), '--thre-hmp', str(THRE_HMP), '--person-thre', str(PERSON_THRE),
                           '--dist-max', str(DIST_MAX)])
    dargs.headnets, dargs.strides, dargs.batch_size = ['hmp', 'omp'], [4, 4], B
    dargs.include_scale = dargs.include_jitter_offset = False
    post = decoder.decoder_factory(dargs)
    hmp_np, omp_np = lowres_inputs(5000 * (rank + 1), B, E, flip)
    hmp_h = torch.from_numpy(hmp_np).pin_memory()
    omp_h = torch.from_numpy(omp_np).pin_memory()

    # ---- the call evaluate.py makes: DEVICE-resident network-resolution maps (right after
    #      model(images)), fused flip + x4 resize + NMS, three batches in flight
    hmp_d, omp_d = hmp_h.to(dev), omp_h.to(dev)
    tables = (cfg.heatmap_hflip(cfg.COCO_KEYPOINTS),) + tuple(cfg.offset_hflip(cfg.COCO_KEYPOINTS, skel))
    tables = tables if flip else None
    for _ in range(args.warmup):
        eng.decode_features(hmp_d, omp_d, 4, 4, 'bicubic', tables)
    barrier()
    dev_l0 = eng.launch_count
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_stages = []
    f0.record()
    depth = min(3, args.steps)            # calls in flight: this path is short enough for the host to matter
    for _ in range(depth - 1):
        eng.decode_features(hmp_d, omp_d, 4, 4, 'bicubic', tables, fetch=False)
    for it in range(args.steps - (depth - 1)):
        eng.decode_features(hmp_d, omp_d, 4, 4, 'bicubic', tables, fetch=False)
        eng.fetch(B)
        if it >= args.steps - depth - 4:  # reading the stage events costs host time this loop does not have
            dev_stages.append(eng.last_stage_times_ms())
    for _ in range(depth - 1):
        eng.fetch(B)
    dev_stages.append(eng.last_stage_times_ms())
    f1.record()
    barrier()
    dev_ms = f0.elapsed_time(f1)
    dev_launches = eng.launch_count - dev_l0

    # ---- baseline leg (N = 1 only, reported inside cpu_baseline): the same decode with stock
    #      PyTorch operators on this GPU (eager ATen kernels for flip fusion, resize, NMS, top-K and
    #      limbs, limbs to the host, multi-threaded CPU grouping) — what the reference's own
    #      formulation costs on the same silicon (oracle/torch_eager.py)
    eager_block = None
    if rank == 0 and world == 1 and not os.environ.get('OG_BENCH_SKIP_EAGER'):      # like cpu_baseline: N = 1 only
        try:
            from oracle import torch_eager
            eager_kw = dict(topk=TOPK, thre_hmp=THRE_HMP, min_len=MIN_LEN, person_thre=PERSON_THRE,
                            dist_max=DIST_MAX, use_scale=True, stride=4, resize_mode='bicubic', flip_test=flip,
                            kp_flips=tables[0] if flip else None, limb_flips=tables[1] if flip else None,
                            limb_reserve=tables[2] if flip else None)
            with torch.no_grad():
                eager_out = torch_eager.generate_poses(hmp_d, omp_d, skel, 17, **eager_kw)      # warm-up
                torch.cuda.synchronize(dev)
                eager_steps = max(2, args.steps // 5)
                t0 = time.perf_counter()
                for _ in range(eager_steps):
                    eager_out = torch_eager.generate_poses(hmp_d, omp_d, skel, 17, **eager_kw)
                torch.cuda.synchronize(dev)
                eager_ms = (time.perf_counter() - t0) * 1e3 / eager_steps
            eager_block = {'value': B / (eager_ms * 1e-3), 'unit': UNIT, 'ms_per_step': eager_ms,
                           'steps': eager_steps, 'persons_per_step': sum(len(p) for p in eager_out),
                           'note': 'stock PyTorch operators on the same GPU, device-resident network-resolution '
                                   'maps (compare with features_dev), grouping on the host cores with the C '
                                   'oracle; one GPU (rank 0), not part of the timed arms'}
            del eager_out
            torch.cuda.empty_cache()
        except Exception as exc:                      # a baseline must never break the bench line
            eager_block = {'unavailable': repr(exc)[:200]}
    del hmp_d, omp_d
    feats = [[[hmp_h], [[]], [[]]], [[omp_h], [[]], [[]]]]
    e2e_eng = post._engine(dev)
    e2e_eng.enable_stage_timing(True)
    for _ in range(args.warmup):
        out = post.generate_poses(feats, flip_test=flip)
    e2e_l0 = e2e_eng.launch_count
    barrier()
    t0 = time.perf_counter()
    e2e_stages = []
    for _ in range(args.steps):
        out = post.generate_poses(feats, flip_test=flip)
        e2e_stages.append(e2e_eng.last_stage_times_ms())
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()
    e2e_launches = e2e_eng.launch_count - e2e_l0
    d2h = sum(p.nbytes for p in out) + (2 * B + 1) * 4
    # host -> device bytes of one step: the heat maps are copied; the offset maps stay in pinned
    # host memory and K2 reads its bilinear samples over PCIe (counted from the candidates:
    # 2 components x 4 taps x 4 bytes per from-candidate of every limb, twice where the mirrored
    # map is averaged in)
    zero_copy = e2e_eng.zero_copy_count > 0
    _, det_idx, _ = e2e_eng.last_intermediates(B)
    per_joint = (det_idx >= 0).sum(dim=2).cpu().numpy()                     # (B, C)
    _, reserved = cfg.offset_hflip(cfg.COCO_KEYPOINTS, skel)
    gathered = 0
    for l, (jf, _jt) in enumerate(skel):
        maps = 2 if (flip and l not in reserved) else 1
        gathered += int(per_joint[:, jf].sum()) * 2 * 4 * 4 * maps
    h2d_full = hmp_h.numel() * 4 + omp_h.numel() * 4
    h2d = hmp_h.numel() * 4 + gathered if zero_copy else h2d_full

    # the same call with the offset maps copied as a whole (zero-copy off), for comparison
    e2e_eng.set_zero_copy(False)
    for _ in range(2):
        post.generate_poses(feats, flip_test=flip)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    full_steps = max(3, args.steps // 2)
    for _ in range(full_steps):
        post.generate_poses(feats, flip_test=flip)
    torch.cuda.synchronize(dev)
    full_copy_ms = (time.perf_counter() - t0) * 1e3 / full_steps
    e2e_launches_all = e2e_eng.launch_count - e2e_l0
    e2e_eng.set_zero_copy(True)

    # ---- max over ranks
    times = torch.tensor([hot_ms, e2e_s * 1e3, dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    hot_ms, e2e_ms, dev_ms = [float(v) for v in times.cpu()]

    if rank == 0:
        k1_ms = statistics.mean(s['k1_stream'] + s['k1_select'] for s in stages)
        k1_stream_ms = statistics.mean(s['k1_stream'] for s in stages)
        stage_mean = {k: statistics.mean(s[k] for s in stages) for k in stages[0]}
        alg_bytes = B * 17 * E * E * 4 + B * 17 * TOPK * 8
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
        try:
            mp = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
            peak, peak_src = float(mp['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
        achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'k1_traffic.json')))['dram_bytes_per_launch']
        except Exception:
            pass
        cpu_block = None          # timed on rank 0 at N = 1 only (bench contract)
        if n_gpus == 1:
            run, cores, what = cpu_decode_fn()
            cpu_n = args.cpu_sample or (16 if cores > 1 else 1)
            cpu_value, cores, what, cpu_times, cpu_stage = time_cpu(args, cpu_n, 2 if cores > 1 else 1)
            cpu_block = {'value': cpu_value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d images of the same workload (flip fusion + x4 resize + NMS/top-K '
                                   '+ limbs + grouping), best of %d, %s' % (cpu_n, len(cpu_times), what),
                         'stage_ms_per_image': cpu_stage,
                         # second baseline of the same leg: the path with stock PyTorch operators on this GPU
                         'eager_torch_gpu': eager_block}
        line = {
            'metric': METRIC, 'value': n_gpus * B * args.steps / (hot_ms * 1e-3), 'unit': UNIT,
            'n_gpus': n_gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': hot_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, n_gpus),
            'e2e': {'value': n_gpus * B * args.steps / (e2e_ms * 1e-3), 'unit': UNIT,
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': e2e_ms / args.steps,
                    'stage_ms': {k: statistics.mean(s[k] for s in e2e_stages) for k in e2e_stages[0]},
                    'fused_redos': e2e_eng.fused_redo_count,
                    'api': 'decoder_factory(args).generate_poses(features, flip_test=%s), pinned host maps' % flip,
                    'h2d_detail': {'heat_maps_copied': hmp_h.numel() * 4,
                                   'offset_samples_read_over_pcie': gathered if zero_copy else 0,
                                   'offset_maps_left_on_host': omp_h.numel() * 4 if zero_copy else 0,
                                   'note': 'fused path: K2 needs 2*L*K bilinear samples per image, so pinned '
                                           'offset maps are read in place (zero-copy) instead of copied'},
                    'full_copy': {'value': n_gpus * B / (full_copy_ms * 1e-3), 'unit': UNIT,
                                  'ms_per_step': full_copy_ms, 'h2d_bytes_per_step': h2d_full,
                                  'note': 'same call with og_set_zero_copy(0): heat and offset maps both copied '
                                          '(rank 0 time, %d steps)' % full_steps}},
            'features_dev': {'value': n_gpus * B * args.steps / (dev_ms * 1e-3), 'unit': UNIT,
                             'ms_per_step': dev_ms / args.steps,
                             'stage_ms': {k: statistics.mean(s[k] for s in dev_stages) for k in dev_stages[0]},
                             'note': 'same API on DEVICE-resident network-resolution maps (what evaluate.py '
                                     'hands over after model(images)): fused flip + x4 bicubic + NMS (K1f), '
                                     'offsets sampled at the candidates; 223 MB of heat maps read per step, '
                                     'no full-resolution map written; three batches in flight'},
            'gpu_launches': hot_launches + e2e_launches_all + dev_launches,
            'gpu_launches_detail': {'hot_path': hot_launches, 'e2e': e2e_launches,
                                    'e2e_full_copy': e2e_launches_all - e2e_launches,
                                    'features_dev': dev_launches},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                         'kernel': 'K1 = nms_candidates_kernel + select_topk_kernel',
                         'algorithmic_bytes_per_launch': alg_bytes, 'ms_per_launch': k1_ms,
                         'stream_pass_only_gbs': alg_bytes / (k1_stream_ms * 1e-3) / 1e9},
            'stage_ms': stage_mean,
            'persons_per_step': n_persons,
            'cpu_baseline': cpu_block,
            'clocks': clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
