"""Shared helpers of the parity tests: error metrics and a log of the measured values."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_l2(a, b):
    import torch
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rel_l2_centered(a, b):
    """Relative L2 error after removing the mean over the batch (dim 0) from both sides: the error measured against the
    part of the signal that differs between samples.  A random-init tower's outputs are dominated by a common vector, so the
    plain relative error would also accept an encoder that ignored its input."""
    import torch
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    a, b = a - a.mean(0, keepdim=True), b - b.mean(0, keepdim=True)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def record_metric(name, **values):
    """Appends a line to gpurun_out/test_metrics.jsonl (brought back from the GPU box) so that measured errors can be read
    next to the bounds the tests assert."""
    import json
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "test_metrics.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=name, **values)) + "\n")
    except OSError:
        pass
