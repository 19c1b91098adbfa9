"""CPU: the host-side cell anchors (paa_b200.anchor_generator.generate_cell_anchors, a restatement of
anchor_generator.py:252-330) against anchors recorded from the reference's generator, and against the live
reference when its tree is mounted."""
import numpy as np
import pytest

from oracle import make_golden, ref_shim
from paa_b200.anchor_generator import generate_cell_anchors
from paa_b200.config import default_cfg
from tests.helpers import load_golden


@pytest.mark.parametrize("case", make_golden.ANCHOR_CASES, ids=[c[0] for c in make_golden.ANCHOR_CASES])
def test_cell_anchors_match_recorded_reference(case):
    name, over, padded, sizes = case
    gold = load_golden("anchors")
    paa = default_cfg(**over).MODEL.PAA
    for l, (stride, size) in enumerate(zip(paa.ANCHOR_STRIDES, paa.ANCHOR_SIZES)):
        per_layer = tuple(paa.OCTAVE ** (k / float(paa.SCALES_PER_OCTAVE)) * size
                          for k in range(paa.SCALES_PER_OCTAVE))
        cell = generate_cell_anchors(stride, per_layer, paa.ASPECT_RATIOS).astype(np.float32)
        a = cell.shape[0]
        assert a == len(paa.ASPECT_RATIOS) * paa.SCALES_PER_OCTAVE
        # location (0, 0) of the recorded grid carries the cell anchors unshifted
        np.testing.assert_array_equal(cell, gold["%s_l%d" % (name, l)][:a])


def test_cell_anchors_match_live_reference():
    if not ref_shim.reference_available():
        pytest.skip("reference tree not mounted")
    ref_shim.load_reference()
    from paa_core.modeling.rpn.anchor_generator import generate_anchors
    for stride, sizes, ratios in [(8, (64,), (1.0,)), (16, (128, 161.27, 203.19), (0.5, 1.0, 2.0)),
                                  (32, (256,), (0.5, 2.0, 3.0)), (128, (1024,), (1.0,)), (4, (32, 40.3), (0.33, 3.0))]:
        np.testing.assert_array_equal(generate_anchors(stride, sizes, ratios).numpy(),
                                      generate_cell_anchors(stride, sizes, ratios))
