"""GPU: each tcgen05 kernel in isolation against torch on exactly bf16-rounded operands (fp32 accumulate),
through the diagnostic C-ABI entry marf_tc_selftest.  Tolerance: fp32 accumulation-order noise only (+ one bf16
rounding of the output for the bf16-storing epilogues)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(mode, rows, K, N, A, W, aux, out_shape):
    from marf_b200 import _lib as L
    lib = L.load()
    out = torch.zeros(out_shape, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    rc = lib.marf_tc_selftest(0, mode, rows, K, N, A.data_ptr(), W.data_ptr() if W is not None else None,
                              aux.data_ptr() if aux is not None else None, out.data_ptr(), C.c_void_p(st))
    assert rc == 0, rc
    return out


def _bf(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("rows,K,N", [(128, 64, 256), (384, 256, 256), (128 * 150, 256, 256), (256, 448, 256), (256, 256, 128)])
def test_forward_bias_relu(rows, K, N):
    torch.manual_seed(rows + K)
    A = torch.randn(rows, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda") * 0.1
    out = _run(0, rows, K, N, A, W, b, (rows, N))
    ref = torch.relu(_bf(A) @ _bf(W).t() + b)
    err = (out - _bf(ref)).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item(), err          # 1 bf16 ulp of the largest value
    assert (out - ref).abs().mean().item() <= 3e-3 * ref.abs().mean().item() + 1e-6


@pytest.mark.parametrize("rows,K,N", [(128, 256, 256), (128 * 149, 256, 256), (512, 256, 128)])
def test_dx_relu_mask(rows, K, N):
    torch.manual_seed(rows)
    A = torch.randn(rows, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    X = torch.relu(torch.randn(rows, N, device="cuda"))
    out = _run(1, rows, K, N, A, W, X, (rows, N))
    ref = (_bf(A) @ _bf(W).t()) * (_bf(X) > 0)
    assert ((out == 0) == (ref == 0)).float().mean().item() > 0.999
    assert (out - _bf(ref)).abs().max().item() <= 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("rows,K", [(128, 256), (128 * 37, 256), (128 * 300, 128)])
def test_dx0_plain_f32(rows, K):
    torch.manual_seed(rows)
    A = torch.randn(rows, K, device="cuda")
    W = torch.randn(64, K, device="cuda") / K ** 0.5
    out = _run(2, rows, K, 64, A, W, None, (rows, 64))
    ref = _bf(A) @ _bf(W).t()
    assert (out - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-5


@pytest.mark.parametrize("rows,K,N", [(128, 256, 256), (128 * 40, 256, 256), (128 * 300, 64, 256), (1280, 448, 256), (640, 256, 128)])
def test_dw(rows, K, N):
    torch.manual_seed(rows + N)
    dY = torch.randn(rows, N, device="cuda")
    X = torch.randn(rows, K, device="cuda")
    out = _run(3, rows, K, N, dY, None, X, (N * K + N,))
    ref = _bf(dY).t() @ _bf(X)
    assert (out[:N * K].view(N, K) - ref).abs().max().item() <= 2e-4 * ref.abs().max().item() + 1e-4
    # fused bias gradient: column sums of dY accumulated by the epilogue warps from the SMEM stages
    bref = _bf(dY).sum(0)
    assert (out[N * K:] - bref).abs().max().item() <= 2e-4 * bref.abs().max().item() + 1e-3
