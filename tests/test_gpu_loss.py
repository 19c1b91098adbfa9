"""GPU parity of the CUDA assign+loss path (through the C ABI) against the oracle and the golden
vectors recorded from the reference.  Stage-wise (teacher-forced) first, end-to-end second
(SURVEY.md 8c).  Tolerances: masks / indices bit-exact apart from the counted tie exemptions;
floats 1e-4 relative (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import make_golden, paa_oracle
from paa_b200 import synthetic
from tests.helpers import (assert_grads_close, check_losses_and_grads_against_oracle, flat_levels, gmm_tie_exempt,
                           load_golden, loss_case_batch, to_device_inputs, topk_tie_exempt)

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _evaluator(**kw):
    import paa_b200
    cfg = paa_b200.default_cfg(**kw)
    return paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))


def _run(ev, batch, requires_grad=True, use_iou=True):
    cls, reg, iou, targets, anchors = to_device_inputs(batch, requires_grad=requires_grad)
    losses = ev(cls, reg, iou if use_iou else None, targets, anchors, None)
    if requires_grad:
        sum(losses).backward()
    torch.cuda.synchronize()
    return losses, cls, reg, iou


def _per_gt_positive_sets(labels, matched, n_gt):
    return [set(np.nonzero((labels > 0) & (matched == g))[0].tolist()) for g in range(n_gt)]


@pytest.mark.parametrize("name", [c[0] for c in make_golden.LOSS_CASES])
def test_stagewise_against_recorded_reference(name):
    ref = load_golden(name)
    b = loss_case_batch(name)
    ev = _evaluator()
    ev.debug = True
    # teacher-forced: selection + GMM consume the reference's own anchor scores
    ev.teacher_combined_loss = torch.from_numpy(ref["combined_loss"]).cuda()
    _run(ev, b, requires_grad=False)
    d = ev.last_debug
    assert np.array_equal(d["matched_idx"].cpu().numpy().astype(np.int64), ref["matched_idx"])
    assert np.array_equal(d["iou_labels"].cpu().numpy(), ref["iou_labels"])
    pos = ref["iou_labels"] > 0
    np.testing.assert_allclose(d["combined_loss"].cpu().numpy()[pos], ref["combined_loss"][pos], rtol=RTOL)
    # oracle run gives the per-GT records for the exemption protocol
    _, _, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred, b.gt_boxes, b.gt_labels,
                                           b.anchors, with_grad=False)
    exempt = topk_tie_exempt(asg) | gmm_tie_exempt(asg)
    got = d["paa_labels"].cpu().numpy()
    total = 0
    for i in range(b.num_images):
        for g in range(b.gt_boxes[i].shape[0]):
            total += 1
            if (i, g) in exempt:
                continue
            sel = ref["matched_idx"][i] == g
            assert np.array_equal(got[i][sel], ref["paa_labels"][i][sel]), (name, i, g)
    assert len(exempt) <= max(1, total // 10)
    # GMM parameters of the fits (teacher-forced => same inputs): iteration counts equal, params 1e-4
    gmm = d["gmm"].cpu().numpy()
    cnt = d["cand_cnt"].cpu().numpy()
    fits = gmm[cnt > 1]
    assert fits.shape[0] == ref["gmm_n"].shape[0]
    assert np.array_equal(cnt[cnt > 1], ref["gmm_n"])
    assert np.array_equal(fits[:, 6].astype(np.int64), ref["gmm_n_iter"])
    np.testing.assert_allclose(fits[:, 0:2], ref["gmm_w"], rtol=RTOL)
    np.testing.assert_allclose(fits[:, 2:4], ref["gmm_mu"], rtol=RTOL)
    np.testing.assert_allclose(fits[:, 4:6], ref["gmm_var"], rtol=RTOL)


@pytest.mark.parametrize("name", [c[0] for c in make_golden.LOSS_CASES])
def test_end_to_end_losses_and_grads_against_recorded_reference(name):
    ref = load_golden(name)
    b = loss_case_batch(name)
    ev = _evaluator()
    ev.debug = True
    losses, cls, reg, iou = _run(ev, b)
    got_labels = ev.last_debug["paa_labels"].cpu().numpy()
    same_labels = np.array_equal(got_labels, ref["paa_labels"])
    _, _, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred, b.gt_boxes,
                                           b.gt_labels, b.anchors, with_grad=False)
    if same_labels:
        np.testing.assert_allclose([float(x) for x in losses], ref["losses"], rtol=RTOL)
        assert_grads_close(flat_levels([t.grad for t in cls]), ref["grad_cls"], rtol=RTOL, atol=1e-9, what="grad_cls")
        assert_grads_close(flat_levels([t.grad for t in reg]), ref["grad_reg"], rtol=RTOL, row_atol=1e-5,
                           what="grad_reg")
        assert_grads_close(flat_levels([t.grad for t in iou])[..., 0], ref["grad_iou"], rtol=RTOL,
                           atol=2e-6 * float(np.abs(ref["grad_iou"]).max()), what="grad_iou")
    else:          # a documented tie flipped a positive set: the flipped GTs must be exempt ...
        exempt = topk_tie_exempt(asg, rel=1e-4) | gmm_tie_exempt(asg, abs_tol=1e-4)
        diff = {(i, int(ref["matched_idx"][i][a])) for i, a in zip(*np.nonzero(got_labels != ref["paa_labels"]))}
        assert diff <= exempt, (name, diff - exempt)
    # ... and in every case losses and gradients must be the oracle's for the labels the device chose
    check_losses_and_grads_against_oracle(b, asg, got_labels, losses, cls, reg, iou)


def _assert_matches_oracle(b, max_exempt, **cfg_overrides):
    """Free-running comparison of one batch with the oracle run on this machine's CPU."""
    names = {"LOSS_GAMMA": "gamma", "LOSS_ALPHA": "alpha"}
    oracle_kw = {names.get(k, k.lower()): v for k, v in cfg_overrides.items()}
    ref_losses, _, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred,
                                                    b.gt_boxes, b.gt_labels, b.anchors,
                                                    params=paa_oracle.default_params(**oracle_kw), with_grad=False)
    ev = _evaluator(**cfg_overrides)
    ev.debug = True
    losses, cls, reg, iou = _run(ev, b)
    d = ev.last_debug
    assert np.array_equal(d["matched_idx"].cpu().numpy().astype(np.int64), asg.matched_idx.numpy())
    pos = asg.iou_labels.numpy() > 0
    np.testing.assert_allclose(d["combined_loss"].cpu().numpy()[pos], asg.combined_loss.numpy()[pos], rtol=RTOL)
    got = d["paa_labels"].cpu().numpy()
    exempt = topk_tie_exempt(asg, rel=1e-4) | gmm_tie_exempt(asg, abs_tol=1e-4)
    diff = {(i, int(asg.matched_idx[i][a])) for i, a in zip(*np.nonzero(got != asg.paa_labels.numpy()))}
    assert diff <= exempt, diff - exempt
    assert len(diff) <= max_exempt
    # never skipped: with flipped (exempt) ties the oracle's stage 5 runs on the device's labels
    check_losses_and_grads_against_oracle(b, asg, got, losses, cls, reg, iou)
    if not diff:
        np.testing.assert_allclose([float(x) for x in losses], [float(x) for x in ref_losses], rtol=RTOL)
    return len(diff)


def test_c1_against_oracle_full_resolution():
    """Config C1 (2 x 800x1333, 20 GT/img) against the oracle run on this machine's CPU."""
    b = synthetic.make_batch(seed=1000, num_images=2, image_hw=(800, 1333), gt_per_image=20)
    _assert_matches_oracle(b, max_exempt=2)


def test_c2_bench_batch_against_oracle():
    """The exact batch bench.py times (config C2: seed 2000, 16 images of 800x1333, 1..100 GT each -- 861 GTs,
    16 x 22 400 anchors) against the oracle: matched indices bit-exact, labels bit-exact modulo counted tie
    exemptions, anchor scores, losses and all gradients to 1e-4."""
    from bench import C2_BATCH_KW
    b = synthetic.make_batch(**C2_BATCH_KW)
    assert b.num_images == 16 and b.num_anchors == 22400
    _assert_matches_oracle(b, max_exempt=8)


def test_without_iou_pred():
    b = synthetic.make_batch(seed=32, num_images=1, image_hw=(256, 256), gt_per_image=5)
    ref_losses, _, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, None, b.gt_boxes, b.gt_labels,
                                                    b.anchors, params=paa_oracle.default_params(use_iou_pred=False),
                                                    with_grad=False)
    ev = _evaluator(USE_IOU_PRED=False)
    ev.debug = True
    losses, *_ = _run(ev, b, requires_grad=True, use_iou=False)      # _ = [cls, reg, iou]
    assert len(losses) == 2
    got = ev.last_debug["paa_labels"].cpu().numpy()
    exempt = topk_tie_exempt(asg, rel=1e-4) | gmm_tie_exempt(asg, abs_tol=1e-4)
    diff = {(i, int(asg.matched_idx[i][a])) for i, a in zip(*np.nonzero(got != asg.paa_labels.numpy()))}
    assert diff <= exempt and len(diff) <= 1, diff - exempt
    cls, reg = _[0], _[1]
    check_losses_and_grads_against_oracle(b, asg, got, losses, cls, reg, None)
    if not diff:
        np.testing.assert_allclose([float(x) for x in losses], [float(x) for x in ref_losses], rtol=RTOL)


@pytest.mark.parametrize("topk", [3, 20])
def test_other_topk_values_against_oracle(topk):
    """TOPK 3 / 20: at most 15 / 100 candidates per GT, i.e. the one- and four-samples-per-lane variants of the
    fit (TOPK 9 uses two), and a per-level list longer than a third of a warp."""
    b = synthetic.make_batch(seed=70 + topk, num_images=2, image_hw=(416, 512), gt_per_image=(3, 10))
    _assert_matches_oracle(b, max_exempt=2, TOPK=topk)


def test_other_focal_parameters_against_oracle():
    """gamma != 2 takes the generic pow path of every focal kernel."""
    b = synthetic.make_batch(seed=91, num_images=2, image_hw=(320, 416), gt_per_image=(2, 8))
    _assert_matches_oracle(b, max_exempt=2, LOSS_GAMMA=1.5, LOSS_ALPHA=0.4)


def test_single_image_single_gt():
    b = synthetic.make_batch(seed=90, num_images=1, image_hw=(256, 320), gt_per_image=1)
    _assert_matches_oracle(b, max_exempt=1)


def test_c3_dense_crowd_against_oracle():
    """Config C3's shape (1333x1333 -> 37 606 anchors, 500 GT) on one image against the oracle."""
    b = synthetic.make_batch(seed=3000, num_images=1, image_hw=(1333, 1333), gt_per_image=500)
    assert b.num_anchors == 37606
    _assert_matches_oracle(b, max_exempt=8)


def test_c5_multiscale_against_oracle():
    """Config C5's shape: images of different sizes padded to the batch maximum (per-image BoxList.size)."""
    hw = synthetic.multiscale_hw(5000, 3)
    b = synthetic.make_batch(seed=5000, num_images=3, image_hw=(0, 0), gt_per_image=(1, 30), per_image_hw=hw)
    assert len(set(b.image_sizes)) > 1
    _assert_matches_oracle(b, max_exempt=2)


@pytest.mark.parametrize("topk", [3, 9, 20])
def test_two_launch_selection_equals_fused(monkeypatch, topk):
    """Calls sized for thousands of GTs run the per-GT program as two launches (selection, then the fit with one
    warp per GT: select_gmm_kernel MODE 1 / 2); it must give bit for bit what the fused block-per-GT form gives --
    labels, per-GT fit parameters, losses and gradients -- here forced on a small batch, and against the oracle."""
    b = synthetic.make_batch(seed=333 + topk, num_images=3, image_hw=(416, 512), gt_per_image=(2, 25))
    outs = []
    # (the same bulk kernel behind both forms: the dynamic one only runs behind the fused form and sums the
    # classification loss in fixed point -- test_bulk_pass_in_the_em_shadow_equals_the_static_pass)
    monkeypatch.setenv("PAA_BULK_EARLY_PCT", "-1")
    for split_above in ("1000000", "0"):
        monkeypatch.setenv("PAA_GMM_SPLIT_ABOVE", split_above)
        ev = _evaluator(TOPK=topk)
        ev.debug = True
        losses, cls, reg, iou = _run(ev, b)
        d = ev.last_debug
        outs.append((d, [float(x) for x in losses], [t.grad.clone() for t in cls + reg + iou]))
    (d0, l0, g0), (d1, l1, g1) = outs
    for k in ("matched_idx", "paa_labels", "cand_idx", "cand_cnt", "num_pos", "gmm", "normalisers"):
        assert torch.equal(d0[k], d1[k]), k
    assert l0 == l1
    for a, c in zip(g0, g1):
        assert torch.equal(a, c)
    monkeypatch.setenv("PAA_GMM_SPLIT_ABOVE", "0")
    _assert_matches_oracle(b, max_exempt=2, TOPK=topk)


def test_bulk_pass_in_the_em_shadow_equals_the_static_pass(monkeypatch):
    """bulk_focal_early_kernel (chunks handed out dynamically; some processed unscaled while the last EM fits run and
    scaled in place afterwards) against the static bulk_focal_kernel: every gradient bit for bit, the regression / IoU
    losses bit for bit, the classification loss to float rounding (it is summed in fixed point) -- and bit for bit
    among all dynamic settings and from run to run, whatever the split into early and late chunks was."""
    b = synthetic.make_batch(seed=4242, num_images=4, image_hw=(800, 1333), gt_per_image=(1, 100))
    outs = {}
    monkeypatch.setenv("PAA_BULK_EARLY_MIN_CHUNKS", "0")        # (calls of this size take the static kernel by default)
    for pct in ("-1", "0", "60", "100", "60"):
        monkeypatch.setenv("PAA_BULK_EARLY_PCT", pct)
        ev = _evaluator()
        losses, cls, reg, iou = _run(ev, b)
        out = ([float(x) for x in losses], [t.grad.clone() for t in cls + reg + iou])
        if pct in outs:
            assert out[0] == outs[pct][0]
        outs.setdefault(pct, out)
    l_static, g_static = outs["-1"]
    for pct in ("0", "60", "100"):
        l_dyn, g_dyn = outs[pct]
        for a, c in zip(g_static, g_dyn):
            assert torch.equal(a, c), pct
        assert l_dyn[1:] == l_static[1:]
        np.testing.assert_allclose(l_dyn[0], l_static[0], rtol=2e-6)
        assert l_dyn == outs["0"][0]
    monkeypatch.setenv("PAA_BULK_EARLY_PCT", "60")
    _assert_matches_oracle(b, max_exempt=2)


@pytest.mark.parametrize("variant", ["default", "no_iou_pred", "gamma", "channels_last", "one_image_one_gt"])
def test_dynamic_bulk_kernel_on_small_calls_equals_static(monkeypatch, variant):
    """Small calls take the static bulk kernel by default; forced onto bulk_focal_early_kernel (fewer chunks than
    blocks, partial last chunks of every level, levels without a whole chunk) they must give the same gradients bit
    for bit and the same losses (classification loss to float rounding), in every configuration the kernel's
    arguments depend on."""
    kw, use_iou, channels_last = {}, True, False
    shape = dict(seed=515, num_images=3, image_hw=(416, 544), gt_per_image=(1, 12))
    if variant == "no_iou_pred":
        kw, use_iou = dict(USE_IOU_PRED=False), False
    elif variant == "gamma":
        kw = dict(LOSS_GAMMA=1.5, LOSS_ALPHA=0.4)
    elif variant == "channels_last":
        channels_last = True
    elif variant == "one_image_one_gt":
        shape = dict(seed=516, num_images=1, image_hw=(256, 320), gt_per_image=1)
    b = synthetic.make_batch(**shape)
    outs = []
    for min_chunks, pct in (("1000000000", "60"), ("0", "60"), ("0", "100"), ("0", "0")):
        monkeypatch.setenv("PAA_BULK_EARLY_MIN_CHUNKS", min_chunks)
        monkeypatch.setenv("PAA_BULK_EARLY_PCT", pct)
        ev = _evaluator(**kw)
        cls, reg, iou, targets, anchors = to_device_inputs(b, requires_grad=True, channels_last=channels_last)
        losses = ev(cls, reg, iou if use_iou else None, targets, anchors, None)
        sum(losses).backward()
        torch.cuda.synchronize()
        grads = [t.grad.clone() for t in cls + reg + (iou if use_iou else [])]
        outs.append(([float(x) for x in losses], grads))
    l_static, g_static = outs[0]
    for l_dyn, g_dyn in outs[1:]:
        for a, c in zip(g_static, g_dyn):
            assert torch.equal(a, c)
        assert l_dyn[1:] == l_static[1:]
        np.testing.assert_allclose(l_dyn[0], l_static[0], rtol=2e-6)
        assert l_dyn == outs[1][0]


def test_dynamic_bulk_kernel_on_a_large_call_equals_static(monkeypatch):
    """More anchors than four per thread of the early bulk kernel's grid (30 images of 800x1333: 672 000 anchors against
    4 x 592 x 256 threads): the tail loop of its zero fill, many chunks per block, positives spread over more than two
    trips of positive_list_kernel's 296 blocks.  Gradients bit for bit, losses as in the small-call test."""
    b = synthetic.make_batch(seed=606, num_images=30, image_hw=(800, 1333), gt_per_image=(60, 100))
    outs = []
    for pct in ("-1", "60"):
        monkeypatch.setenv("PAA_BULK_EARLY_PCT", pct)
        ev = _evaluator()
        losses, cls, reg, iou = _run(ev, b)
        outs.append(([float(x) for x in losses], [t.grad.clone() for t in cls + reg + iou]))
        del cls, reg, iou
    (l_static, g_static), (l_dyn, g_dyn) = outs
    for a, c in zip(g_static, g_dyn):
        assert torch.equal(a, c)
    assert l_dyn[1:] == l_static[1:]
    np.testing.assert_allclose(l_dyn[0], l_static[0], rtol=2e-6)


def _check_full_size_properties(b):
    """Size-independent checks: determinism, positives are candidates matched to their GT, every GT with
    candidates gets >= 1 positive, and the batch splits into halves with identical labels and additive
    normalisers."""
    ev = _evaluator()
    ev.debug = True
    cls, reg, iou, targets, anchors = to_device_inputs(b)
    l1 = [float(x) for x in ev(cls, reg, iou, targets, anchors, None)]
    d1 = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in ev.last_debug.items()}
    l2 = [float(x) for x in ev(cls, reg, iou, targets, anchors, None)]
    d2 = ev.last_debug
    assert l1 == l2
    for k in ("matched_idx", "paa_labels", "cand_idx", "num_pos", "gmm"):
        assert torch.equal(d1[k], d2[k]), k
    labels = d1["paa_labels"].cpu().numpy()
    matched = d1["matched_idx"].cpu().numpy()
    cand = d1["cand_idx"].cpu().numpy()
    cnt = d1["cand_cnt"].cpu().numpy()
    npos = d1["num_pos"].cpu().numpy()
    off = d1["gt_offsets"]
    assert int((labels > 0).sum()) == int(npos.sum()) == int(round(float(d1["normalisers"][0])))
    for i in range(b.num_images):
        for g in range(off[i + 1] - off[i]):
            gi = off[i] + g
            c = cand[gi, :cnt[gi]]
            assert (matched[i][c] == g).all()
            assert set(np.nonzero((labels[i] > 0) & (matched[i] == g))[0]) == set(c[:npos[gi]])
            assert npos[gi] >= (1 if cnt[gi] > 0 else 0)
            assert (labels[i][c[:npos[gi]]] == int(b.gt_labels[i][g])).all()
    # halves
    half = b.num_images // 2
    norm_sum = np.zeros(2)
    for sl in (slice(0, half), slice(half, None)):
        ev(  # noqa
            [t[sl] for t in cls], [t[sl] for t in reg], [t[sl] for t in iou], targets[sl], anchors[sl], None)
        dh = ev.last_debug
        assert torch.equal(dh["paa_labels"], d1["paa_labels"][sl])
        norm_sum += dh["normalisers"].cpu().numpy()
    np.testing.assert_allclose(norm_sum, d1["normalisers"].cpu().numpy(), rtol=1e-12)


def test_full_size_properties_c2():
    """The bench shape: 16 images of 800x1333, 1..100 GT each."""
    _check_full_size_properties(synthetic.make_batch(seed=2000, num_images=16, image_hw=(800, 1333),
                                                     gt_per_image=(1, 100)))


def test_full_size_properties_c3():
    """Dense crowd, one rank's share: 4 images of 1333x1333 with 500 GT each."""
    _check_full_size_properties(synthetic.make_batch(seed=3001, num_images=4, image_hw=(1333, 1333),
                                                     gt_per_image=500))


def test_full_size_properties_c5():
    """Multi-scale, one rank's share: 8 images of different sizes padded to the batch maximum."""
    hw = synthetic.multiscale_hw(5001, 8)
    _check_full_size_properties(synthetic.make_batch(seed=5001, num_images=8, image_hw=(0, 0),
                                                     gt_per_image=(1, 100), per_image_hw=hw))


def test_error_behaviour():
    import paa_b200
    b = synthetic.make_batch(seed=5, num_images=2, image_hw=(128, 160), gt_per_image=3)
    cls, reg, iou, targets, anchors = to_device_inputs(b)
    ev = _evaluator()
    empty = paa_b200.BoxList(torch.zeros((0, 4), device="cuda"), targets[0].size)
    empty.add_field("labels", torch.zeros(0, dtype=torch.int64, device="cuda"))
    with pytest.raises(ValueError):          # matcher.py:53-58
        ev(cls, reg, iou, [empty, targets[1]], anchors, None)
    wrong = paa_b200.BoxList(targets[0].bbox, (999, 999))
    wrong.add_field("labels", targets[0].get_field("labels"))
    with pytest.raises(RuntimeError):        # boxlist_ops.py:95-97
        ev(cls, reg, iou, [wrong, targets[1]], anchors, None)
    with pytest.raises(RuntimeError):        # no CPU path
        ev([t.cpu() for t in cls], reg, iou, targets, anchors, None)


def test_upstream_gradient_scaling():
    b = synthetic.make_batch(seed=6, num_images=1, image_hw=(160, 160), gt_per_image=3)
    ev = _evaluator()
    cls, reg, iou, targets, anchors = to_device_inputs(b, requires_grad=True)
    sum(ev(cls, reg, iou, targets, anchors, None)).backward()
    g1 = [t.grad.clone() for t in cls + reg + iou]
    for t in cls + reg + iou:
        t.grad = None
    l = ev(cls, reg, iou, targets, anchors, None)
    (2.0 * l[0] + 3.0 * l[1] + 0.5 * l[2]).backward()
    L = len(cls)
    for k, t in enumerate(cls + reg + iou):
        w = (2.0, 3.0, 0.5)[k // L]
        torch.testing.assert_close(t.grad, g1[k] * w, rtol=1e-6, atol=1e-12)
    # gradients of one call taken twice (retain_graph) with different upstream weights
    heads = cls + reg + iou
    l = ev(cls, reg, iou, targets, anchors, None)
    first = [g.clone() for g in torch.autograd.grad(2.0 * l[0] + 3.0 * l[1] + 0.5 * l[2], heads, retain_graph=True)]
    second = torch.autograd.grad(l[0] + 0.25 * l[1] + 4.0 * l[2], heads)
    for k in range(len(heads)):
        torch.testing.assert_close(first[k], g1[k] * (2.0, 3.0, 0.5)[k // L], rtol=1e-6, atol=1e-12)
        torch.testing.assert_close(second[k], g1[k] * (1.0, 0.25, 4.0)[k // L], rtol=2e-6, atol=1e-12)


def test_candidate_pool_overflow_falls_back_to_tile_scan(monkeypatch):
    """A (GT, level) pool that overflows is re-collected by scanning the tiles: identical labels."""
    b = synthetic.make_batch(seed=9, num_images=2, image_hw=(416, 512), gt_per_image=(3, 9))
    ev = _evaluator()
    ev.debug = True
    _run(ev, b, requires_grad=False)
    want = {k: ev.last_debug[k].clone() for k in ("paa_labels", "cand_idx", "num_pos")}
    assert int(ev.last_debug["cand_cnt"].max()) > 8
    monkeypatch.setenv("PAA_SEG_CAP", "4")
    _run(ev, b, requires_grad=False)
    for k, v in want.items():
        assert torch.equal(ev.last_debug[k], v), k
