"""rlobjectdetection_b200 -- B200-native (sm_100a) detection hot path
proposal -> NMS -> RoIAlign / RoIPool -> RL-refine, behind the module surface of
jbr97/RLObjectDetection (`model.rpn.proposal_layer._ProposalLayer`, `model.nms.nms_wrapper.nms`,
`model.roi_align.modules.roi_align.RoIAlignAvg`, `model.roi_pooling.modules.roi_pool._RoIPooling`,
`model.rpn.bbox_transform.*`, `model.Reinforcement.action.Action`).

Everything computes in hand-written CUDA kernels reached through the C ABI of
csrc/librlod_sm100a.so (include/rlod.h); there is no CPU or eager-PyTorch fallback.

Drop-in use under the reference's drivers: put this directory on sys.path *before* the
reference's lib/ so that `import model.rpn.proposal_layer` resolves here.
"""
from . import model  # noqa: F401
from .model import _backend as backend  # noqa: F401

__all__ = ["model", "backend"]
