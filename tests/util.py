"""Shared helpers for the parity tests (test infrastructure, not product code)."""
import os

import torch

BF16 = torch.bfloat16


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def assert_close(name: str, got: torch.Tensor, want: torch.Tensor, tol: float):
    """Relative L2 error bound `tol` (stated per test) plus finiteness and shape checks."""
    assert tuple(got.shape) == tuple(want.shape), f"{name}: shape {tuple(got.shape)} != {tuple(want.shape)}"
    assert torch.isfinite(got.float()).all(), f"{name}: non-finite values"
    err = rel_l2(got, want)
    path = os.environ.get("MMGL_PARITY_REPORT")
    if path:
        with open(path, "a") as f:
            f.write(f"{os.environ.get('PYTEST_CURRENT_TEST', '?')}\n  {name:36s} rel-L2 {err:.3e} (tol {tol:.1e})\n")
    assert err <= tol, f"{name}: rel-L2 {err:.3e} > {tol:.1e} (max|d| {max_abs(got, want):.3e})"


def bf16r(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 and back to fp32: the oracle then sees exactly the values the kernel sees."""
    return t.to(BF16).float()


def randn(gen, *shape, scale=1.0, device="cuda"):
    return (torch.randn(*shape, generator=gen) * scale).to(device)


class Report:
    """Collect every comparison of a test, print them all, then fail once (so one GPU run shows the whole picture)."""

    def __init__(self):
        self.rows, self.bad = [], []

    def close(self, name, got, want, tol):
        ok_shape = tuple(got.shape) == tuple(want.shape)
        finite = bool(torch.isfinite(got.float()).all())
        err = rel_l2(got, want) if ok_shape else float("inf")
        self.rows.append(f"{name:36s} rel-L2 {err:.3e} (tol {tol:.1e}) max|d| {max_abs(got, want) if ok_shape else -1:.3e}")
        if not (ok_shape and finite and err <= tol):
            self.bad.append(name)

    def scalar(self, name, got, want, rtol, atol):
        got, want = float(got), float(want)
        self.rows.append(f"{name:36s} got {got:.6g} want {want:.6g} (rtol {rtol:.1e} atol {atol:.1e})")
        if not abs(got - want) <= rtol * abs(want) + atol:
            self.bad.append(name)

    def absolute(self, name, got, want, atol):
        err = max_abs(got, want)
        self.rows.append(f"{name:36s} max|d| {err:.3e} (atol {atol:.1e})")
        if not err <= atol:
            self.bad.append(name)

    def finish(self):
        print("\n".join(self.rows))
        path = os.environ.get("MMGL_PARITY_REPORT")      # one GPU run -> the measured error of every comparison
        if path:
            with open(path, "a") as f:
                f.write(os.environ.get("PYTEST_CURRENT_TEST", "?") + "\n  " + "\n  ".join(self.rows) + "\n")
        assert not self.bad, "out of tolerance: " + ", ".join(self.bad)
