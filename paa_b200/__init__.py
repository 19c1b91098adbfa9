"""paa_b200 -- the PAA anchor-assignment / loss / post-processing hot path as hand-written sm_100a
CUDA kernels behind a C ABI, with the reference's Python call signatures on top.

Drop-in seam: ``make_paa_loss_evaluator`` and ``make_paa_postprocessor`` replace
the factories of paa_core/modeling/rpn/paa/{loss,inference}.py that ``PAAModule.__init__`` calls
(paa.py:117-119).
"""
from paa_b200.box_coder import BoxCoder
from paa_b200.config import default_cfg
from paa_b200.structures import BoxList, boxlist_iou, cat_boxlist

__all__ = ["BoxCoder", "BoxList", "boxlist_iou", "cat_boxlist", "default_cfg", "make_paa_loss_evaluator",
           "make_paa_postprocessor", "PAALossComputation", "PAAPostProcessor", "make_anchor_generator_paa",
           "AnchorGenerator", "make_atss_postprocessor", "ATSSPostProcessor", "make_retinanet_postprocessor",
           "RetinaNetPostProcessor", "make_fcos_postprocessor", "FCOSPostProcessor", "make_atss_loss_evaluator",
           "ATSSLossComputation", "make_retinanet_loss_evaluator", "RetinaNetLossComputation",
           "make_fcos_loss_evaluator", "FCOSLossComputation", "make_rpn_loss_evaluator", "RPNLossComputation",
           "BalancedPositiveNegativeSampler"]


def __getattr__(name):
    # the evaluator classes load libpaa_b200.so; import them lazily so that CPU-only tooling
    # (synthetic inputs, config) does not need the library
    if name in ("make_paa_loss_evaluator", "PAALossComputation", "make_atss_loss_evaluator", "ATSSLossComputation",
                "make_retinanet_loss_evaluator", "RetinaNetLossComputation", "make_fcos_loss_evaluator",
                "FCOSLossComputation"):
        from paa_b200 import loss
        return getattr(loss, name)
    if name in ("make_paa_postprocessor", "PAAPostProcessor", "make_atss_postprocessor", "ATSSPostProcessor",
                "make_retinanet_postprocessor", "RetinaNetPostProcessor", "make_fcos_postprocessor",
                "FCOSPostProcessor"):
        from paa_b200 import inference
        return getattr(inference, name)
    if name in ("make_rpn_loss_evaluator", "RPNLossComputation", "BalancedPositiveNegativeSampler"):
        from paa_b200 import rpn_loss
        return getattr(rpn_loss, name)
    if name in ("make_anchor_generator_paa", "AnchorGenerator"):
        from paa_b200 import anchor_generator
        return getattr(anchor_generator, name)
    raise AttributeError(name)
