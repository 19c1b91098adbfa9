"""BoxCoder handle with the reference's constructor (modeling/rpn/atss/atss.py:14-17).

PAA builds ``BoxCoder(cfg)`` and hands it to both factories (paa.py:117-119).  In this framework the
encode / decode arithmetic (atss.py:33-50, 68-96; weights 10,10,5,5, clamp log(1000/16)) lives inside
the CUDA kernels, so the object only carries the configuration and validates that it is the
variant the kernels implement.
"""


class BoxCoder(object):
    WEIGHTS = (10.0, 10.0, 5.0, 5.0)

    def __init__(self, cfg):
        self.cfg = cfg
        self.regression_type = cfg.MODEL.ATSS.REGRESSION_TYPE

    def check_supported(self):
        if self.regression_type != "BOX":
            raise NotImplementedError(
                "paa_b200 implements the 'BOX' regression type used by every PAA config "
                "(defaults.py:367); got %r" % (self.regression_type,))


def coder_regression_type(box_coder):
    """Accepts this BoxCoder or the reference's (which keeps the cfg)."""
    if hasattr(box_coder, "regression_type"):
        return box_coder.regression_type
    return box_coder.cfg.MODEL.ATSS.REGRESSION_TYPE
