"""Drop-in replacement of the reference's ``decoder`` package: the public names of the
reference's ``decoder/__init__.py`` with the same signatures, backed by the sm_100a CUDA
library through ``offsetguided_b200._lib``."""
from . import collect, factory, group, heatmap, offset

# stand-alone stages
hmp_NMS = heatmap.hmp_NMS
topK_channel = heatmap.topK_channel
joint_dets = heatmap.joint_dets
scored_offset = offset.scored_offset
soft_nms = group.soft_nms
# classes and the factory evaluate.py / demo_batch.py use
LimbsCollect = collect.LimbsCollect
GreedyGroup = group.GreedyGroup
PostProcess = factory.PostProcess
decoder_cli = factory.decoder_cli
decoder_factory = factory.decoder_factory

__all__ = ['hmp_NMS', 'topK_channel', 'joint_dets', 'scored_offset', 'soft_nms', 'LimbsCollect',
           'GreedyGroup', 'PostProcess', 'decoder_cli', 'decoder_factory']
