"""Drop-in replacement of the reference's ``decoder`` package
(reference decoder/__init__.py:1-5): same names, same signatures, backed by the
sm_100a CUDA library through ``offsetguided_b200._lib``."""
from .heatmap import hmp_NMS, topK_channel, joint_dets
from .offset import scored_offset
from .group import GreedyGroup, soft_nms
from .collect import LimbsCollect
from .factory import decoder_factory, decoder_cli, PostProcess
