"""K1 alone (og_nms_topk_f32) and decode_maps without / with a second call in flight, on the bench's
full-resolution maps: which part of the hot-path K1 time is the kernel and which is interference."""
import os
import sys
import json

ROOT = os.environ.get('OG_TREE') or os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch            # noqa: E402
from offsetguided_b200 import config as cfg          # noqa: E402
from offsetguided_b200.engine import DecoderEngine   # noqa: E402
from oracle import scenes                            # noqa: E402

skel = cfg.COCO_PERSON_SKELETON
h8, o8 = scenes.synth_hires_batch(1000, 8, 6, 640, 640, skel)
heat = torch.from_numpy(h8).cuda().repeat(8, 1, 1, 1).contiguous()
offs = torch.from_numpy(o8).cuda().repeat(8, 1, 1, 1).contiguous()
eng = DecoderEngine(17, skel, topk=32, thre_hmp=0.04, min_len=0.5, dist_max=40, use_scale=True, person_thre=0.04)
for _ in range(10):
    eng.nms_topk(heat)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for _ in range(20):
    eng.nms_topk(heat)
b.record()
torch.cuda.synchronize()
out = {'tree': ROOT, 'k1_alone_ms': a.elapsed_time(b) / 20}
eng.enable_stage_timing(True)
for _ in range(10):
    eng.decode_maps(heat, offs)
st = []
for _ in range(10):
    eng.decode_maps(heat, offs)
    st.append(eng.last_stage_times_ms()['k1_stream'])
out['k1_stream_depth1_ms'] = sum(st) / len(st)
st = []
eng.decode_maps(heat, offs, fetch=False)
for _ in range(20):
    eng.decode_maps(heat, offs, fetch=False)
    eng.fetch(64)
    st.append(eng.last_stage_times_ms()['k1_stream'])
eng.fetch(64)
out['k1_stream_depth2_ms'] = sum(st) / len(st)
print(json.dumps(out))
