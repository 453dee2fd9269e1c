"""CPU checks of bench.py's synthetic workload (no GPU, no kernels): the training batches hold 8-bit image values k/255, so the
resident step, the host-buffer step fed with fp32 arrays, the one fed with bytes and the reference arm all see the same numbers;
the JSON `config` object is the same on both arms."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_training_batches_are_8bit_valued_and_round_trip_through_bytes():
    w = dict(bench.WORKLOADS["tiny"], B=3)
    y, g, h = bench.make_inputs(w, 1001)
    for t in (y, g):
        assert t.dtype == torch.float32 and float(t.min()) >= 0.0 and float(t.max()) <= 1.0
        u8 = (t * 255).round().to(torch.uint8)                      # what e2e_train uploads
        assert torch.equal(u8.float() / 255, t)                      # exactly the values the float paths consume
    assert h.shape[-1] == w["k"] and abs(float(h.sum()) - 1.0) < 1e-5


def test_forward_workloads_keep_fp32_samples_and_config_is_arm_independent():
    w = dict(bench.WORKLOADS["cfg5"], B=2)
    y, _, _ = bench.make_inputs(w, 7)
    assert not torch.equal((y * 255).round() / 255, y)               # inference workloads: unquantised blurred + noisy images
    c = bench.config_of(bench.WORKLOADS["cfg2_train"], "cfg2_train")
    assert c["mode"] == "fwd+bwd" and c["iters"] == 10 and "8-bit" in c["samples"] and c["per_gpu_batch"] == 64
