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
