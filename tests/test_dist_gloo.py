"""World-size-2 check of the path's multi-rank protocol on CPU (gloo).

The CUDA kernels cannot run here, so the per-rank kernel stages are played by the oracle; what is
under test is the product's host-side exchange (`paa_b200.loss.reduce_normalisers`, WORLD_SIZE
semantics of loss.py:18-28) and the identity it must satisfy: with images sharded over W ranks and
the two normalisers summed across ranks, the mean over ranks of every loss equals the single-process
loss on the whole batch (what DDP's gradient averaging relies on)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import paa_oracle
from paa_b200 import synthetic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from paa_b200.loss import get_num_gpus, reduce_normalisers
        assert get_num_gpus() == world
        b = synthetic.make_batch(seed=51, num_images=4, image_hw=(192, 256), gt_per_image=(2, 7))
        per = b.num_images // world
        sl = slice(rank * per, (rank + 1) * per)
        heads = ([t[sl] for t in b.box_cls], [t[sl] for t in b.box_regression], [t[sl] for t in b.iou_pred])
        asg = paa_oracle.assign(*heads, b.gt_boxes[sl], b.gt_labels[sl], b.anchors)
        norm = torch.tensor([float(asg.num_pos), asg.sum_iou], dtype=torch.float64)
        local = norm.clone()
        reduce_normalisers(norm)                                   # the product's exchange step
        losses = paa_oracle.losses(*heads, asg, total_num_pos=float(norm[0]), total_sum_iou=float(norm[1]),
                                   world_size=world)
        np.save(os.path.join(out_dir, "rank%d.npy" % rank),
                np.array([float(x) for x in losses] + local.tolist() + norm.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_normaliser_exchange_matches_single_process(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / ("rank%d.npy" % k)) for k in range(world)]
    # both ranks saw the same totals, equal to the sum of the locals
    np.testing.assert_allclose(r[0][5:7], r[1][5:7], rtol=0, atol=0)
    np.testing.assert_allclose(r[0][5:7], r[0][3:5] + r[1][3:5], rtol=1e-12)
    # single process on the whole batch
    b = synthetic.make_batch(seed=51, num_images=4, image_hw=(192, 256), gt_per_image=(2, 7))
    ref, _, asg = paa_oracle.assign_and_loss(b.box_cls, b.box_regression, b.iou_pred, b.gt_boxes, b.gt_labels,
                                             b.anchors, with_grad=False)
    assert asg.num_pos == int(round(r[0][5]))
    mean_over_ranks = (r[0][:3] + r[1][:3]) / world
    np.testing.assert_allclose(mean_over_ranks, [float(x) for x in ref], rtol=1e-5)


def test_single_process_is_a_noop(monkeypatch):
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    from paa_b200.loss import get_num_gpus, reduce_normalisers
    assert get_num_gpus() == 1
    t = torch.tensor([3.0, 1.5], dtype=torch.float64)
    assert reduce_normalisers(t) is t and t.tolist() == [3.0, 1.5]


# ---- the reference itself under WORLD_SIZE = 2 (only where /root/reference is mounted) -----------------------
def _rank_share(b, rank, world):
    import dataclasses
    per = b.num_images // world
    sl = slice(rank * per, (rank + 1) * per)
    return dataclasses.replace(b, image_sizes=b.image_sizes[sl], gt_boxes=b.gt_boxes[sl], gt_labels=b.gt_labels[sl],
                               box_cls=[t[sl] for t in b.box_cls], box_regression=[t[sl] for t in b.box_regression],
                               iou_pred=[t[sl] for t in b.iou_pred])


def _reference_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import make_golden
        b = synthetic.make_batch(seed=52, num_images=4, image_hw=(224, 288), gt_per_image=(2, 9))
        ref = make_golden.run_reference_loss(_rank_share(b, rank, world))      # loss.py:18-28,321,338 over gloo
        np.savez(os.path.join(out_dir, "ref_rank%d.npz" % rank), losses=ref["losses"], grad_cls=ref["grad_cls"],
                 grad_reg=ref["grad_reg"], grad_iou=ref["grad_iou"], paa_labels=ref["paa_labels"])
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_reference_under_two_ranks_matches_the_oracles_normalisation(tmp_path):
    """The unmodified reference evaluator run by two gloo ranks on two halves of a batch (its own `reduce_sum`,
    loss.py:22-28, averaging by WORLD_SIZE, loss.py:322,338) against the oracle's stage 5 with the summed
    normalisers -- the rule the CUDA path's exchange implements (`PaaLossArgs.world_size`)."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not mounted")
    from tests.helpers import flat_levels
    world = 2
    mp.spawn(_reference_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    b = synthetic.make_batch(seed=52, num_images=4, image_hw=(224, 288), gt_per_image=(2, 9))
    shares = [_rank_share(b, r, world) for r in range(world)]
    asgs = [paa_oracle.assign(s.box_cls, s.box_regression, s.iou_pred, s.gt_boxes, s.gt_labels, s.anchors)
            for s in shares]
    total_pos = float(sum(a.num_pos for a in asgs))
    total_iou = float(sum(a.sum_iou for a in asgs))
    assert asgs[0].num_pos != asgs[1].num_pos            # the two ranks really normalise by a shared total
    for r in range(world):
        ref = np.load(tmp_path / ("ref_rank%d.npz" % r))
        s = shares[r]
        heads = ([t.clone().requires_grad_(True) for t in s.box_cls],
                 [t.clone().requires_grad_(True) for t in s.box_regression],
                 [t.clone().requires_grad_(True) for t in s.iou_pred])
        assert np.array_equal(asgs[r].paa_labels.numpy(), ref["paa_labels"])
        losses = paa_oracle.losses(*heads, asgs[r], total_num_pos=total_pos, total_sum_iou=total_iou, world_size=world)
        sum(losses).backward()
        np.testing.assert_allclose([float(x.detach()) for x in losses], ref["losses"], rtol=1e-6)
        np.testing.assert_allclose(flat_levels([t.grad for t in heads[0]]), ref["grad_cls"], rtol=1e-5, atol=1e-10)
        np.testing.assert_allclose(flat_levels([t.grad for t in heads[1]]), ref["grad_reg"], rtol=1e-5, atol=1e-10)


def _flavour_worker(rank, world, port, out_dir):
    """The reference's ATSS and FCOS evaluators (atss/loss.py:14-23,255-268; fcos/loss.py:22-31,250-279) on this
    rank's half of a batch, normalisers summed over gloo."""
    import types
    import warnings
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        warnings.simplefilter("ignore")
        from oracle import ref_shim
        from tests.test_oracle_fcos_vs_reference import fcos_batch
        ref = ref_shim.load_reference()
        from paa_core.modeling.rpn.atss import loss as aloss
        from paa_core.modeling.rpn.fcos import loss as floss
        ns = types.SimpleNamespace
        out = {}
        # ATSS
        b = _rank_share(synthetic.make_batch(seed=53, num_images=4, image_hw=(384, 512), gt_per_image=(2, 9)), rank, world)
        cfg = ns(MODEL=ns(ATSS=ns(LOSS_GAMMA=(2.0,), LOSS_ALPHA=(0.25,), FG_IOU_THRESHOLD=0.5, BG_IOU_THRESHOLD=0.4,
                                  POSITIVE_TYPE="ATSS", TOPK=9, REG_LOSS_WEIGHT=2.0, REGRESSION_TYPE="BOX")))
        targets = []
        for i in range(b.num_images):
            t = ref.BoxList(b.gt_boxes[i], b.image_sizes[i])
            t.add_field("labels", b.gt_labels[i])
            targets.append(t)
        anchors = [[ref.BoxList(a, b.image_sizes[i]) for a in b.anchors] for i in range(b.num_images)]
        with torch.no_grad():
            rl = aloss.ATSSLossComputation(cfg, ref.BoxCoder(cfg))(b.box_cls, b.box_regression, b.iou_pred, targets,
                                                                    anchors)
        out["atss"] = np.array([float(x) for x in rl])
        # FCOS
        fb, locations = fcos_batch(54, (384, 512), (2, 9), num_images=4)
        fb = _rank_share(fb, rank, world)
        cfg = ns(MODEL=ns(FCOS=ns(LOSS_GAMMA=(2.0,), LOSS_ALPHA=(0.25,), FPN_STRIDES=[8, 16, 32, 64, 128],
                                  CENTER_SAMPLING_RADIUS=1.5, IOU_LOSS_TYPE="giou", NORM_REG_TARGETS=True)))
        targets = []
        for i in range(fb.num_images):
            t = ref.BoxList(fb.gt_boxes[i], fb.image_sizes[i])
            t.add_field("labels", fb.gt_labels[i])
            targets.append(t)
        with torch.no_grad():
            rl = floss.make_fcos_loss_evaluator(cfg)(locations, fb.box_cls, fb.box_regression, fb.iou_pred, targets)
        out["fcos"] = np.array([float(x) for x in rl])
        np.savez(os.path.join(out_dir, "flavours_rank%d.npz" % rank), **out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_reference_atss_and_fcos_under_two_ranks_match_the_oracles_normalisation(tmp_path):
    from oracle import atss_oracle, fcos_oracle, ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not mounted")
    from tests.test_oracle_fcos_vs_reference import fcos_batch
    world = 2
    mp.spawn(_flavour_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / ("flavours_rank%d.npz" % r)) for r in range(world)]
    # ATSS
    b = synthetic.make_batch(seed=53, num_images=4, image_hw=(384, 512), gt_per_image=(2, 9))
    shares = [_rank_share(b, r, world) for r in range(world)]
    asgs = [atss_oracle.assign(s.gt_boxes, s.gt_labels, s.anchors, atss_oracle.default_params()) for s in shares]
    tot_pos, tot_ctr = float(sum(a.num_pos for a in asgs)), float(sum(a.sum_centerness for a in asgs))
    for r, s in enumerate(shares):
        with torch.no_grad():
            ol = atss_oracle.losses(s.box_cls, s.box_regression, s.iou_pred, asgs[r], total_num_pos=tot_pos,
                                    total_sum_centerness=tot_ctr, world_size=world)
        np.testing.assert_allclose([float(x) for x in ol], got[r]["atss"], rtol=1e-6)
    # FCOS
    fb, locations = fcos_batch(54, (384, 512), (2, 9), num_images=4)
    prm = fcos_oracle.default_params(center_sampling_radius=1.5, iou_loss_type="giou", norm_reg_targets=True)
    shares = [_rank_share(fb, r, world) for r in range(world)]
    asgs = [fcos_oracle.assign(s.gt_boxes, s.gt_labels, locations, prm) for s in shares]
    tot_pos, tot_ctr = float(sum(a.num_pos for a in asgs)), float(sum(a.sum_centerness for a in asgs))
    for r, s in enumerate(shares):
        with torch.no_grad():
            ol = fcos_oracle.losses(s.box_cls, s.box_regression, s.iou_pred, asgs[r], total_num_pos=tot_pos,
                                    total_sum_centerness=tot_ctr, world_size=world)
        np.testing.assert_allclose([float(x) for x in ol], got[r]["fcos"], rtol=1e-6)


# ---- the peer exchange's set-up falls back collectively, before anybody enters the rendezvous ---------------------
def _fallback_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from paa_b200 import loss as paa_loss

        def rendezvous_must_not_run(self, device, buffer):
            raise AssertionError("rank %d entered the rendezvous although a peer could not allocate" % rank)

        paa_loss.PeerNormExchange.__init__ = rendezvous_must_not_run
        if rank == 0:      # this rank could allocate; rank 1 runs the real allocation, which fails on a CPU device
            paa_loss.PeerNormExchange._allocate = staticmethod(lambda device: torch.zeros(8, dtype=torch.float64))
        import warnings
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            state = paa_loss.PeerNormExchange.get(torch.device("cpu"))
        assert state is None
        # only the rank that failed explains why; both go on to the all-reduce
        assert (len([w for w in caught if "peer-memory" in str(w.message)]) > 0) == (rank == 1)
        t = torch.tensor([1.0 + rank, 2.0], dtype=torch.float64)
        paa_loss.reduce_normalisers(t)
        np.save(os.path.join(out_dir, "fallback%d.npy" % rank), t.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_peer_exchange_falls_back_collectively_when_one_rank_cannot_allocate(tmp_path):
    world = 2
    mp.spawn(_fallback_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for k in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / ("fallback%d.npy" % k)), [3.0, 4.0])
