"""Deterministic synthetic inputs for parity tests, smoke() and bench.py.

All draws use a CPU ``torch.Generator`` so that the CUDA path and the CPU oracle see
bit-identical inputs.  Images get one bright Gaussian blob each so that the per-image
maximum of the blurred sensor image is well separated from the runner-up: the
reference's ``amax`` normalisation (``Face-DeId/Camera/Optics.py:128``) back-propagates
through the arg-max, and a near-tie flips it between FFT implementations (trap T2).
"""
from __future__ import annotations

import torch

__all__ = ["height_map", "images", "upstream_grad", "top2_relative_gap"]


def height_map(N: int, seed: int = 1234, amplitude: float = 1e-6) -> torch.Tensor:
    """(1,N,N) fp32 lens height in metres, uniform in +-amplitude."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand(1, N, N, generator=g) * 2 - 1) * amplitude


def images(B: int, N: int, seed: int = 1000, channels: int = 3) -> torch.Tensor:
    """(B,C,N,N) fp32 in [0,1]: uniform noise + one Gaussian blob per image, max-normalised."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    img = torch.rand(B, channels, N, N, generator=g)
    margin = min(40, N // 4)
    centres = margin + torch.rand(B, 2, generator=g) * (N - 2 * margin)
    ax = torch.arange(N, dtype=torch.float32)
    sigma = 12.0 * N / 256.0
    gy = torch.exp(-0.5 * ((ax[None, :] - centres[:, 0:1]) / sigma) ** 2)   # (B,N)
    gx = torch.exp(-0.5 * ((ax[None, :] - centres[:, 1:2]) / sigma) ** 2)
    blob = 2.0 * gy[:, None, :, None] * gx[:, None, None, :]
    img = img + blob
    return img / img.amax((1, 2, 3), keepdim=True)


def upstream_grad(B: int, N: int, seed: int = 2000, channels: int = 3) -> torch.Tensor:
    """(B,C,N,N) fp32 weights w for the synthetic loss sum(y*w)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.rand(B, channels, N, N, generator=g)


def top2_relative_gap(conv: torch.Tensor) -> torch.Tensor:
    """Per-image (max - runner-up)/max of a (B,C,H,W) tensor."""
    flat = conv.reshape(conv.shape[0], -1)
    top = flat.topk(2, dim=1).values
    return (top[:, 0] - top[:, 1]) / top[:, 0]
