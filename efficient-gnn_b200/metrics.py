"""Calibration metrics on the device: accuracy, mean confidence and the
class-wise ECE the reference reports (utils/ece.py:8-89, evaluated as
benchmark_calibration_methods.py:100-127 does), without moving the [N,C]
probabilities to the host.  Thin call into libegnn_b200 (no CPU fallback)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi

__all__ = ["calibration_metrics"]


def calibration_metrics(outputs: torch.Tensor, labels: torch.Tensor, mask=None, *, logits: bool = False,
                        log_probs: bool = True, n_bins: int = 10):
    """``(accuracy, mean max-probability, class-wise ECE)`` of ``outputs``
    ``[N,C]`` on the samples selected by the boolean ``mask``.

    ``outputs`` are log-probabilities by default (what ``WATS.forward``
    returns); ``log_probs=False`` means probabilities, ``logits=True`` applies a
    softmax first (the reference's ``calculate_average_ece(logits=True)``)."""
    _cabi.require_device()
    lib = _cabi.load()
    x = outputs.detach()
    if not x.is_cuda:
        x = x.cuda()
    if logits:
        x, log_probs = torch.log_softmax(x.float(), dim=1), True
    x = x.to(torch.float32).contiguous()
    n, c = x.shape
    y = labels.detach().to(device=x.device, dtype=torch.int64).contiguous()
    m = None if mask is None else mask.detach().to(device=x.device).to(torch.uint8).contiguous()
    with torch.cuda.device(x.device):
        out = torch.empty(3, dtype=torch.float64, device=x.device)
        ws_bytes = int(lib.egnn_calibration_metrics_ws_bytes(c, n_bins))
        if ws_bytes == 0:
            raise ValueError("n_bins must be in [1, 32]")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        _cabi.check(lib.egnn_calibration_metrics(
            _cabi.ptr(x), 1 if log_probs else 0, _cabi.ptr(y), _cabi.ptr(m), n, c, n_bins, _cabi.ptr(out),
            _cabi.ptr(ws), ws_bytes, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
            "egnn_calibration_metrics")
    acc, conf, ece = out.cpu().tolist()
    return acc, conf, ece
