"""The peer-memory exchange of the loss normalisers when a rank is late or never arrives: the wait
has no deadline by default (here: one far beyond the delay) -- a contribution that arrives seconds late is used as if
it had been on time -- and with a deadline a missing contribution is an ERROR (status record + trap), never NaN data.  One GPU is enough: the second
rank is played by a side stream that writes (or does not write) its slot of this rank's exchange buffer, which is
exactly what the peer's kernel does over NVLink.  Runs in a child process: the trap poisons the CUDA context."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, time
sys.path.insert(0, %(root)r)
os.environ["WORLD_SIZE"] = "2"
import numpy as np, torch
import paa_b200
from paa_b200 import _lib, loss as paa_loss, synthetic
from paa_b200.synthetic import to_device_inputs

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)

class FakePeers(object):
    """Two "ranks" whose exchange buffers both live on this GPU; this process is rank 0."""
    def __init__(self, timeout_s):
        self.world, self.rank = 2, 0
        self.bufs = [torch.zeros(_lib.PEER_BUFFER_DOUBLES, dtype=torch.float64, device=dev) for _ in range(2)]
        self.ptrs = [b.data_ptr() for b in self.bufs]
        self.status = torch.zeros(4, dtype=torch.int32).pin_memory()
        self._status_np = self.status.numpy()
        self.timeout_s = timeout_s
    raise_if_timed_out = paa_loss.PeerNormExchange.raise_if_timed_out

def run(mode):
    # "late": a deadline far beyond the delay (the product default is none at all; a test must never leave a kernel
    # spinning forever on a shared GPU should the "peer" copy fail to run)
    timeout_s = 20.0 if mode == "late" else 0.3
    peers = FakePeers(timeout_s)
    cfg = paa_b200.default_cfg()
    ev = paa_b200.make_paa_loss_evaluator(cfg, paa_b200.BoxCoder(cfg))
    ev.debug = True
    b = synthetic.make_batch(seed=41, num_images=2, image_hw=(256, 320), gt_per_image=(2, 6))
    cls, reg, iou, targets, anchors = to_device_inputs(b)
    side = torch.cuda.Stream()
    vals = torch.tensor([5.0, 2.5], dtype=torch.float64, device=dev)
    one = torch.ones(1, dtype=torch.float64, device=dev)
    scratch = torch.zeros(8, dtype=torch.float64, device=dev)
    # Load every kernel of the step (and the copies the "peer" will issue) BEFORE a kernel spins on this GPU: CUDA
    # loads modules lazily at first launch, and a load synchronises the context -- it would wait for the spinning
    # kernel, i.e. for data that only this very thread can deliver.  (With real peers the data comes from another
    # GPU, so the first step merely serialises.)
    os.environ["WORLD_SIZE"] = "1"
    ev.forward_backward(cls, reg, iou, targets, anchors)
    with torch.cuda.stream(side):
        scratch[0:2].copy_(vals)
        scratch[2:3].copy_(one)
    torch.cuda.synchronize()
    os.environ["WORLD_SIZE"] = "2"
    paa_loss.PeerNormExchange.get = classmethod(lambda cls, device: peers)
    t0 = time.time()
    losses, grads = ev.forward_backward(cls, reg, iou, targets, anchors)       # returns at once: all stream work
    if mode == "late":
        # rank 1's contribution lands 2.5 s late (the old code gave up after ~2 s and produced NaN losses)
        time.sleep(2.5)
        slot = (1 * _lib.MAX_PEERS + 1) * 4          # first step: epoch 1 -> parity 1; rank 1
        with torch.cuda.stream(side):                # stream order: the data, then the epoch that makes it visible
            peers.bufs[0][slot:slot + 2].copy_(vals)
            peers.bufs[0][slot + 2:slot + 3].copy_(one)
        torch.cuda.synchronize()
        waited = time.time() - t0
        d = ev.last_debug
        local, total = d["local_normalisers"].cpu().numpy(), d["normalisers"].cpu().numpy()
        assert waited >= 2.4, waited
        np.testing.assert_array_equal(total, local + np.array([5.0, 2.5]))
        l = losses.cpu().numpy()
        assert np.isfinite(l).all() and (l > 0).all(), l
        assert int(peers.status[0]) == 0
        print("LATE_PEER_OK waited %%.2f s" %% waited)
    else:
        try:
            torch.cuda.synchronize()
        except Exception as e:           # the trap: "unspecified launch failure" on every later call
            print("CUDA_ERROR", type(e).__name__)
        else:
            raise SystemExit("the wait returned although rank 1 never published")
        assert time.time() - t0 >= 0.25
        assert time.time() - t0 < 30.0
        st = peers.status.numpy()
        assert st[0] == 1 and st[1] == 1 and st[2] == 1, st
        try:
            peers.raise_if_timed_out()
        except RuntimeError as e:
            assert "rank 1" in str(e)
            print("TIMEOUT_REPORTED")

run(sys.argv[1])
'''


def _child(mode, tmp_path):
    script = tmp_path / "peer_child.py"
    script.write_text(CHILD % dict(root=ROOT))
    return subprocess.run([sys.executable, str(script), mode], capture_output=True, text=True, timeout=300)


def test_a_late_peer_is_waited_for(tmp_path):
    r = _child("late", tmp_path)
    assert r.returncode == 0 and "LATE_PEER_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


def test_a_missing_peer_is_an_error_not_nan_when_a_deadline_is_set(tmp_path):
    r = _child("never", tmp_path)
    assert "CUDA_ERROR" in r.stdout and "TIMEOUT_REPORTED" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
