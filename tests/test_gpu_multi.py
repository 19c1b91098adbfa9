"""Two-GPU check of the step's only exchange (loss.py:321,338): the peer-memory exchange of the normalisers
gives the same losses / gradients as the NCCL all-reduce and as the oracle's two-rank result, eagerly and
under CUDA-graph replay.  Skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    import paa_b200
    from paa_b200 import loss as paa_loss, synthetic
    from tests.helpers import to_device_inputs
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    b = synthetic.make_batch(seed=700 + rank, num_images=2, image_hw=(320, 416), gt_per_image=(2, 8))
    cfg = paa_b200.default_cfg()
    ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
    ev.debug = True
    cls, reg, iou, targets, anchors = to_device_inputs(b, device=dev)
    res = {}
    peer = paa_loss.PeerNormExchange.get(dev)
    res["peer_available"] = peer is not None
    losses, grads = ev.forward_backward(cls, reg, iou, targets, anchors)
    res["peer_losses"] = losses.cpu().numpy()
    res["peer_norm"] = ev.last_debug["normalisers"].cpu().numpy()
    res["peer_grad0"] = grads["cls"][0].cpu().numpy()
    # a rank that arrives 3 s late (rank 0 saving a checkpoint, a data-loader stall) is simply waited for
    if rank == 1:
        import time
        time.sleep(3.0)
    res["peer_losses_delayed"] = ev.forward_backward(cls, reg, iou, targets, anchors)[0].cpu().numpy()
    # graph replay: the epoch counter lives in device memory, so replays stay in step across ranks
    ev.debug = False
    for _ in range(2):
        ev.forward_backward(cls, reg, iou, targets, anchors)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ev.forward_backward(cls, reg, iou, targets, anchors)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        gl, _ = ev.forward_backward(cls, reg, iou, targets, anchors)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    res["graph_losses"] = gl.cpu().numpy()
    # NCCL path for comparison
    paa_loss.PeerNormExchange._by_device[(dev.type, dev.index)] = None
    losses2, grads2 = ev.forward_backward(cls, reg, iou, targets, anchors)
    res["nccl_losses"] = losses2.cpu().numpy()
    res["nccl_grad0"] = grads2["cls"][0].cpu().numpy()
    # the ATSS flavour publishes its normalisers from another kernel (atss_norm_kernel): same check
    from types import SimpleNamespace as NS
    acfg = NS(MODEL=NS(ATSS=NS(LOSS_GAMMA=2.0, LOSS_ALPHA=0.25, TOPK=9, REG_LOSS_WEIGHT=2.0, POSITIVE_TYPE="ATSS",
                               REGRESSION_TYPE="BOX")))
    ba = synthetic.make_batch(seed=710 + rank, num_images=2, image_hw=(384, 512), gt_per_image=(2, 8))
    acls, areg, actr, atargets, aanchors = to_device_inputs(ba, device=dev)
    aev = paa_b200.make_atss_loss_evaluator(acfg, paa_b200.BoxCoder(acfg))
    res["atss_nccl_losses"] = aev.forward_backward(acls, areg, actr, atargets, aanchors)[0].cpu().numpy()
    paa_loss.PeerNormExchange._by_device.pop((dev.type, dev.index), None)
    assert paa_loss.PeerNormExchange.get(dev) is not None
    res["atss_peer_losses"] = aev.forward_backward(acls, areg, actr, atargets, aanchors)[0].cpu().numpy()
    # FCOS publishes from the same fold kernel; RetinaNet normalises by its own rank's counts (no exchange)
    fcfg = NS(MODEL=NS(FCOS=NS(LOSS_GAMMA=2.0, LOSS_ALPHA=0.25, FPN_STRIDES=[8, 16, 32, 64, 128],
                               CENTER_SAMPLING_RADIUS=1.5, IOU_LOSS_TYPE="giou", NORM_REG_TARGETS=True)))
    bf = synthetic.make_batch(seed=720 + rank, num_images=2, image_hw=(320, 416), gt_per_image=(2, 8),
                              trained_like=False)
    bf.box_regression = [(t.abs() * 40.0 + 1.0) for t in bf.box_regression]
    fcls, freg, fctr, ftargets, _ = to_device_inputs(bf, device=dev)
    flocs = [p.to(dev) for p in synthetic.fcos_locations(bf.grids)]
    fev = paa_b200.make_fcos_loss_evaluator(fcfg)
    res["fcos_peer_losses"] = fev.forward_backward(flocs, fcls, freg, fctr, ftargets)[0].cpu().numpy()
    paa_loss.PeerNormExchange._by_device[(dev.type, dev.index)] = None
    res["fcos_nccl_losses"] = fev.forward_backward(flocs, fcls, freg, fctr, ftargets)[0].cpu().numpy()
    rcfg = NS(MODEL=NS(RETINANET=NS(LOSS_GAMMA=2.0, LOSS_ALPHA=0.25, FG_IOU_THRESHOLD=0.5, BG_IOU_THRESHOLD=0.4,
                                    BBOX_REG_BETA=0.11, BBOX_REG_WEIGHT=4.0)))
    br = synthetic.make_retinanet_batch(seed=730 + rank, num_images=2, image_hw=(320, 416), gt_per_image=(2, 7))
    rcls, rreg, _, rtargets, ranchors = to_device_inputs(br, device=dev)
    rev = paa_b200.make_retinanet_loss_evaluator(rcfg, NS(weights=(10.0, 10.0, 5.0, 5.0)))
    res["retinanet_losses"] = rev.forward_backward(ranchors, rcls, rreg, rtargets)[0].cpu().numpy()
    torch.cuda.synchronize()
    dist.barrier()
    np.savez(out % rank, **res)
    os._exit(0)


def test_peer_exchange_matches_all_reduce_and_oracle(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from oracle import paa_oracle
    from paa_b200 import synthetic
    out = str(tmp_path / "rank%d.npz")
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    res = [dict(np.load(out % r)) for r in range(2)]
    # oracle: two ranks, normalisers summed over ranks (loss.py:321,338 with WORLD_SIZE=2)
    asgs = []
    batches = [synthetic.make_batch(seed=700 + r, num_images=2, image_hw=(320, 416), gt_per_image=(2, 8))
               for r in range(2)]
    for b in batches:
        asgs.append(paa_oracle.assign(b.box_cls, b.box_regression, b.iou_pred, b.gt_boxes, b.gt_labels, b.anchors))
    tot_pos = sum(a.num_pos for a in asgs)
    tot_iou = sum(a.sum_iou for a in asgs)
    for r in range(2):
        assert bool(res[r]["peer_available"])
        np.testing.assert_allclose(res[r]["peer_norm"], [tot_pos, tot_iou], rtol=1e-6)
        np.testing.assert_array_equal(res[r]["peer_losses"], res[r]["nccl_losses"])
        np.testing.assert_array_equal(res[r]["peer_losses"], res[r]["graph_losses"])
        np.testing.assert_array_equal(res[r]["peer_losses"], res[r]["peer_losses_delayed"])
        np.testing.assert_array_equal(res[r]["peer_grad0"], res[r]["nccl_grad0"])
        np.testing.assert_array_equal(res[r]["atss_peer_losses"], res[r]["atss_nccl_losses"])
        assert np.isfinite(res[r]["atss_peer_losses"]).all()
        np.testing.assert_array_equal(res[r]["fcos_peer_losses"], res[r]["fcos_nccl_losses"])
        assert np.isfinite(res[r]["fcos_peer_losses"]).all()
        from oracle import retinanet_oracle
        br = synthetic.make_retinanet_batch(seed=730 + r, num_images=2, image_hw=(320, 416), gt_per_image=(2, 7))
        want, _, _ = retinanet_oracle.assign_and_loss(br.box_cls, br.box_regression, br.gt_boxes, br.gt_labels,
                                                      br.anchors, with_grad=False)
        np.testing.assert_allclose(res[r]["retinanet_losses"][:2], [float(x) for x in want], rtol=1e-4)
        b = batches[r]
        ref = paa_oracle.losses(b.box_cls, b.box_regression, b.iou_pred, asgs[r], total_num_pos=tot_pos,
                                total_sum_iou=tot_iou, world_size=2)
        np.testing.assert_allclose(res[r]["peer_losses"], [float(x) for x in ref[:3]], rtol=1e-4)
