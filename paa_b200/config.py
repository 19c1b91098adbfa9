"""The PAA configuration keys the path reads (paa_core/config/defaults.py:292-331,:548), as a plain
attribute tree.  yacs ``CfgNode`` objects from the reference work unchanged: only attribute access
is used."""
from types import SimpleNamespace


def default_cfg(**paa_overrides):
    paa = dict(NUM_CLASSES=81, ANCHOR_SIZES=(64, 128, 256, 512, 1024), ASPECT_RATIOS=(1.0,),
               ANCHOR_STRIDES=(8, 16, 32, 64, 128), STRADDLE_THRESH=0, OCTAVE=2.0, SCALES_PER_OCTAVE=1,
               LOSS_ALPHA=0.25, LOSS_GAMMA=2.0, IOU_THRESHOLD=0.1, TOPK=9, REG_LOSS_WEIGHT=1.3,
               PRIOR_PROB=0.01, INFERENCE_TH=0.05, NMS_TH=0.6, PRE_NMS_TOP_N=1000, USE_IOU_PRED=True,
               IOU_LOSS_WEIGHT=0.5, INFERENCE_SCORE_VOTING=True, REG_LOSS_TYPE="iou")
    paa.update(paa_overrides)
    ns = SimpleNamespace
    return ns(MODEL=ns(PAA=ns(**paa), ATSS=ns(REGRESSION_TYPE="BOX")),
              TEST=ns(DETECTIONS_PER_IMG=100, BBOX_AUG=ns(ENABLED=False, VOTE=False)))


def scalar(v):
    """LOSS_GAMMA / LOSS_ALPHA are floats in defaults.py but the reference's CPU focal loss indexes
    them (sigmoid_focal_loss.py:42-43), so configs in the wild carry 1-tuples as well."""
    if isinstance(v, (tuple, list)):
        return float(v[0])
    return float(v)
