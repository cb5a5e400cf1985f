"""Shared test utilities: tolerances and dtype handling."""

import torch

EPS = {torch.bfloat16: 2.0**-8, torch.float16: 2.0**-11, torch.float32: 2.0**-16}


def rounded(t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """fp32 copy of ``t`` after rounding to ``dtype`` - what the kernel actually sees."""
    return t.to(dtype).float()


def assert_close(actual: torch.Tensor, expected: torch.Tensor, dtype: torch.dtype, scale: float = 1.0, what: str = "") -> None:
    """|actual - expected| <= scale * eps(dtype) * max|expected| (+ tiny): output rounding plus
    accumulation-order noise, far below what any indexing or protocol bug produces."""
    actual = actual.detach().float().cpu()
    expected = expected.detach().float().cpu()
    assert actual.shape == expected.shape, f"{what}: shape {tuple(actual.shape)} != {tuple(expected.shape)}"
    assert torch.isfinite(actual).all(), f"{what}: non-finite output"
    ref_max = expected.abs().max().item() if expected.numel() else 0.0
    tol = scale * EPS[dtype] * max(ref_max, 1e-3) + 1e-6
    err = (actual - expected).abs().max().item() if expected.numel() else 0.0
    assert err <= tol, f"{what}: max abs err {err:.3e} > tol {tol:.3e} (ref max {ref_max:.3e})"
