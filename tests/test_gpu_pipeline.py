"""The host-fed FramePipeline (copy-in / kernels / copy-out overlapped over several streams and slots) must give exactly
what one device-resident call gives, and what the oracle gives, whatever the batching: ragged last batch, more batches
than slots, float32 or float64 host frames, pinned or pageable input, per-frame histograms, alternating kernel streams.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import port  # noqa: E402  (the checker)
from waterorderlib_b200 import engine, synth  # noqa: E402
from waterorderlib_b200.pipeline import FramePipeline  # noqa: E402


def frames(n_frames, m=4, sigma=0.4):
    xyz, boxes = synth.trajectory(m, n_frames, sigma=sigma, seed0=50)
    return xyz, boxes


@pytest.mark.parametrize("dtype,batch,slots,streams", [(np.float64, 2, 2, 2), (np.float32, 3, 4, 2), (np.float64, 1, 3, 1),
                                                       (np.float32, 16, 2, 3)])
def test_pipeline_matches_device_run_and_oracle(dtype, batch, slots, streams):
    xyz, boxes = frames(11)
    F, N = xyz.shape[:2]
    dev = engine.q3b_frames(torch.from_numpy(xyz).cuda(), boxes, want=("q", "n3", "nn_idx", "ang_hist", "q_hist", "frame_stats"))
    pipe = FramePipeline(N, batch, dtype=dtype, n_slots=slots, n_run_streams=streams, want_nn=True)
    host = torch.from_numpy(xyz.astype(dtype)).pin_memory()
    for rep in range(2):  # a second run reuses the slots and workspaces
        r = pipe.run(host, boxes)  # no synchronize here: run() returns when the results are in host memory
        assert torch.equal(r["q"], dev["q"].cpu()) and torch.equal(r["n3"], dev["n3"].cpu())
        assert torch.equal(r["nn_idx"], dev["nn_idx"].cpu())
        assert torch.equal(r["ang_hist"], dev["ang_hist"].cpu()) and torch.equal(r["q_hist"], dev["q_hist"].cpu())
        assert np.allclose(r["frame_stats"].numpy(), dev["frame_stats"].cpu().numpy(), rtol=1e-12, atol=1e-12)
        assert pipe.h2d_bytes == F * N * 3 * np.dtype(dtype).itemsize and pipe.launches > 0
    for f in (0, F - 1):  # first frame and the ragged last batch against the oracle
        ref = port.three_body(xyz[f], xyz[f], boxes[f])
        assert np.array_equal(r["n3"][f].numpy(), ref["numAngs"])
        assert np.allclose(r["q"][f].numpy(), port.getOrderParamq(xyz[f], xyz[f], boxes[f]), rtol=0, atol=1e-6)
    total = sum(port.three_body(xyz[f], xyz[f], boxes[f])["hist"] for f in range(F))
    assert np.array_equal(r["ang_hist"][0].numpy(), total)


def test_pipeline_per_frame_histograms_pageable_input_and_errors():
    xyz, boxes = frames(5)
    F, N = xyz.shape[:2]
    pipe = FramePipeline(N, 2, dtype=np.float64, hist_per_frame=True)
    r = pipe.run(xyz, boxes)  # numpy, pageable
    torch.cuda.synchronize()
    assert r["ang_hist"].shape == (F, 500)
    for f in range(F):
        assert np.array_equal(r["ang_hist"][f].numpy(), port.three_body(xyz[f], xyz[f], boxes[f])["hist"])
    q_only = FramePipeline(N, 4, dtype=np.float64, do_3body=False)
    rq = q_only.run(xyz, boxes[0])  # one box for all frames
    torch.cuda.synchronize()
    assert rq["n3"] is None and "ang_hist" not in rq and torch.equal(rq["q"], r["q"])
    with pytest.raises(ValueError):
        pipe.run(xyz.astype(np.float32), boxes)
    with pytest.raises(ValueError):
        pipe.run(xyz[:, :-1], boxes)


def test_pipeline_histograms_only_is_what_the_drivers_consume():
    """want_q = want_n3 = False: nothing per water comes back, histograms and per-frame sums are unchanged."""
    xyz, boxes = frames(7)
    F, N = xyz.shape[:2]
    full = FramePipeline(N, 2, dtype=np.float32).run(xyz.astype(np.float32), boxes)
    pipe = FramePipeline(N, 3, dtype=np.float32, want_q=False, want_n3=False)
    r = pipe.run(torch.from_numpy(xyz.astype(np.float32)).pin_memory(), boxes)
    assert r["q"] is None and r["n3"] is None
    assert torch.equal(r["ang_hist"], full["ang_hist"]) and torch.equal(r["q_hist"], full["q_hist"])
    assert np.allclose(r["frame_stats"].numpy(), full["frame_stats"].numpy(), rtol=1e-12, atol=1e-12)
    assert pipe.d2h_bytes == (2 * 500 + F * 8) * 8 and r["n_overflow_last_batches"] == 0


def test_pipeline_timeline_trace():
    xyz, boxes = frames(6)
    pipe = FramePipeline(xyz.shape[1], 2, dtype=np.float64)
    pipe.trace = True
    pipe.run(torch.from_numpy(xyz).pin_memory(), boxes)
    tl = pipe.timeline()
    assert tl.shape == (3, 6) and np.all(tl[:, 1] >= tl[:, 0]) and np.all(tl[:, 3] >= tl[:, 2]) and np.all(tl[:, 5] >= tl[:, 4])
    assert np.all(tl[:, 2] >= tl[:, 1] - 1e-3) and np.all(tl[:, 4] >= tl[:, 3] - 1e-3)  # copy-in -> kernels -> copy-out per batch


def test_pipeline_fp32_arithmetic_mode():
    """precision="fp32" through the pipeline: float results identical to the device-resident fp32 call, and within the
    1e-4 bar of the fp64 oracle wherever both modes picked the same four neighbours."""
    xyz, boxes = frames(5)
    F, N = xyz.shape[:2]
    dev = engine.q3b_frames(torch.from_numpy(xyz).cuda(), boxes, precision="fp32")
    pipe = FramePipeline(N, 2, dtype=np.float32, precision="fp32", want_nn=True)
    r = pipe.run(torch.from_numpy(xyz.astype(np.float32)).pin_memory(), boxes)
    torch.cuda.synchronize()
    assert r["q"].dtype == torch.float32 and torch.equal(r["q"], dev["q"].cpu()) and torch.equal(r["nn_idx"], dev["nn_idx"].cpu())
    assert torch.equal(r["ang_hist"], dev["ang_hist"].cpu()) and torch.equal(r["n3"], dev["n3"].cpu())
    q64, nn64, _ = port.order_param_q(xyz[0], xyz[0], boxes[0], 0.0, 10.0)
    same = np.all(r["nn_idx"][0].numpy() == nn64, axis=1)
    assert same.mean() > 0.99 and np.max(np.abs(r["q"][0].numpy()[same] - q64[same])) < 1e-4
