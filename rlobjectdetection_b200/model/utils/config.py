"""The configuration keys the hot path reads, under the reference's names
(lib/model/utils/config.py:112-119, 141-147, 175, 182, 191-199, 283-298).  Only these keys are provided:
the reference's dataset / solver / yaml machinery is out of scope.  `cfg` is an attribute
dict, so `cfg[cfg_key].RPN_PRE_NMS_TOP_N` (proposal_layer.py:72-75) and `cfg.POOLING_SIZE`
(faster_rcnn.py:33-34) work unchanged; `cfg_from_list` accepts the same flat key/value list as
the reference's --set option."""
import ast


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


cfg = AttrDict(
    TRAIN=AttrDict(RPN_NMS_THRESH=0.7, RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000,
                   RPN_MIN_SIZE=8,
                   # target layers (config.py:76-87, 114, 131-153)
                   BATCH_SIZE=128, FG_FRACTION=0.25, FG_THRESH=0.5, BG_THRESH_HI=0.5, BG_THRESH_LO=0.1,
                   BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0), RPN_POSITIVE_OVERLAP=0.7, RPN_NEGATIVE_OVERLAP=0.3,
                   RPN_CLOBBER_POSITIVES=False, RPN_FG_FRACTION=0.5, RPN_BATCHSIZE=256,
                   RPN_BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0), RPN_POSITIVE_WEIGHT=-1.0,
                   # test_net.py:251-260 un-normalises bbox_pred with these (config.py:112-119)
                   BBOX_NORMALIZE_TARGETS_PRECOMPUTED=True, BBOX_NORMALIZE_MEANS=(0.0, 0.0, 0.0, 0.0),
                   BBOX_NORMALIZE_STDS=(0.1, 0.1, 0.2, 0.2)),
    TEST=AttrDict(NMS=0.3, RPN_NMS_THRESH=0.7, RPN_PRE_NMS_TOP_N=6000, RPN_POST_NMS_TOP_N=300,
                  RPN_MIN_SIZE=16, BBOX_REG=True),
    POOLING_MODE="align",  # the reference defaults to 'crop' (:283); all three modes are provided
    CROP_RESIZE_WITH_MAX_POOL=True,
    POOLING_SIZE=7,
    MAX_NUM_GT_BOXES=20,
    ANCHOR_SCALES=[8, 16, 32],
    ANCHOR_RATIOS=[0.5, 1, 2],
    FEAT_STRIDE=[16],
    CUDA=True,
)


def cfg_from_list(cfg_list):
    """['TEST.RPN_POST_NMS_TOP_N', '100', ...] -> cfg (reference: config.py:376-399)."""
    if len(cfg_list) % 2:
        raise ValueError("cfg_from_list expects key/value pairs")
    for key, value in zip(cfg_list[0::2], cfg_list[1::2]):
        node = cfg
        *path, leaf = key.split(".")
        for part in path:
            node = node[part]
        if leaf not in node:
            raise KeyError(key)
        try:
            value = ast.literal_eval(value) if isinstance(value, str) else value
        except (ValueError, SyntaxError):
            pass
        node[leaf] = value
