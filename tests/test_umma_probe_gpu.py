"""Pins the tcgen05 shared-memory descriptor semantics for row-shifted views of a SWIZZLE_128B tile (developer probe).
The halo-tile convolution design needs: start address shifted by whole 128 B rows (not 1024-aligned) and an SBO
that is not 1024 (8-row groups taken from different image lines of a halo tile)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(A, B, shift, sbo, base_offset, a_mn):
    from passl_b200 import _lib
    from passl_b200.kernels import _ptr, _stream
    lib = _lib.load()
    out = torch.full((128, 64), float("nan"), device="cuda")
    _lib.check(lib.passl_b200_umma_probe(_ptr(A), _ptr(B), _ptr(out), shift, sbo, base_offset, a_mn, _stream()), "umma_probe")
    torch.cuda.synchronize()
    return out


def _expect_kmajor(A, B, shift, sbo):
    m = torch.arange(128, device="cuda")
    rows = shift + (m // 8) * (sbo // 128) + m % 8
    return A.float()[rows] @ B.float().t()


def _expect_mn(A, B, shift):
    k = torch.arange(64, device="cuda")
    lo = A.float()[shift + k].t() @ B.float().t()            # [64 ch, 64 n]
    hi = A.float()[shift + 128 + k].t() @ B.float().t()
    return torch.cat([lo, hi], 0)


def test_probe_report():
    """Not an assertion of one behaviour but a report + the invariant the conv kernels use: at least the aligned case works and
    the shifted cases work with one of the two base_offset conventions; the winning convention is printed."""
    torch.manual_seed(0)
    A = torch.randn(256, 64, device="cuda").bfloat16()
    B = torch.randn(64, 64, device="cuda").bfloat16()
    rep = []
    ok_k = {}
    for shift in (0, 8, 1, 3, 10, 21):
        for sbo in (1024, 1280):
            for bo in (0, shift & 7):
                got = _run(A, B, shift, sbo, bo, 0)
                err = (got - _expect_kmajor(A, B, shift, sbo)).abs().max().item()
                rep.append("K-major shift=%d sbo=%d base_offset=%d err=%.3g" % (shift, sbo, bo, err))
                ok_k[(shift, sbo, bo)] = err < 0.5
    ok_m = {}
    for shift in (0, 8, 1, 3, 10, 21):
        for bo in (0, -1):
            got = _run(A, B, shift, 1024, bo, 1)
            err = (got - _expect_mn(A, B, shift)).abs().max().item()
            rep.append("MN-major shift=%d base_offset=%s err=%.3g" % (shift, "auto" if bo < 0 else "0", err))
            ok_m[(shift, bo)] = err < 0.5
    print("\n".join(rep))
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/umma_probe_report.txt", "w").write("\n".join(rep) + "\n")
    assert ok_k[(0, 1024, 0)] and ok_k[(8, 1024, 0)] and ok_m[(0, 0)] and ok_m[(8, 0)]
