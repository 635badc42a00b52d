"""Helpers shared by the GPU parity tests (tests only; may import the oracle)."""
import numpy as np
import torch


def bf16_grid(a):
    """Round a float32 numpy array to the nearest bf16 value (returned as float32)."""
    return torch.as_tensor(np.asarray(a, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def dev_bf16(a, dev):
    return torch.as_tensor(np.asarray(a, np.float32)).to(torch.bfloat16).to(dev).contiguous()


def dev_f32(a, dev):
    return torch.as_tensor(np.asarray(a, np.float32)).to(dev).contiguous()


def host(t):
    return t.detach().to(torch.float32).cpu().numpy()


def rel_err(got, ref):
    """max |got-ref| / max |ref|  (per-tensor)."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    d = np.abs(got - ref).max()
    m = np.abs(ref).max()
    return float(d / m) if m > 0 else float(d)


def cosine(got, ref):
    a = np.asarray(got, np.float64).ravel()
    b = np.asarray(ref, np.float64).ravel()
    na, nb = np.linalg.norm(a), np.linalg.norm(b)
    if na == 0 and nb == 0:
        return 1.0
    return float(a @ b / (na * nb + 1e-300))


def assert_close(got, ref, tol, what=""):
    e = rel_err(got, ref)
    c = cosine(got, ref)
    assert e <= tol and c >= 0.9999, f"{what}: rel_err {e:.3e} (tol {tol:.1e}) cosine {c:.6f}"
