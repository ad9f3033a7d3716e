"""GPU: the 3xTF32 tensor-core GEMMs of precision=fp32 (csrc/tc_tf32.cuh) in isolation, through the diagnostic C-ABI entry
marf_tf32_gemm, against torch float64.  Tolerance: the error must stay within a small multiple of what an fp32 GEMM
(torch, TF32 off) makes on the same operands — the 3-term split is there to be fp32-grade, not TF32-grade (1e-3)."""
import ctypes as C

import pytest
import torch

import planar_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import gpu_util
    cfg = po.PlanarConfig(H=40, W=56, patch_H=20, patch_W=28, batch_size=2, use_masks=False)
    e = gpu_util.make_engine(cfg, "fp32")
    yield e
    e.close()


def _gemm(eng, mode, epi, M, N, K, A, W, Cout, aux=None):
    from marf_b200 import _lib as L
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.marf_tf32_gemm(eng.handle, mode, epi, M, N, K, A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), Cout.data_ptr(),
                            Cout.stride(0), aux.data_ptr() if aux is not None else None,
                            aux.stride(0) if aux is not None and aux.dim() == 2 else 0, C.c_void_p(st))
    assert rc == 0, (rc, lib.marf_last_error(eng.handle))
    return Cout


def _check(out, ref64, ref32):
    """error vs float64 no worse than 4x the fp32 GEMM's (+ a floor of 2e-6 of the peak)"""
    peak = ref64.abs().max().item()
    e_tc = (out.double() - ref64).abs().max().item()
    e_32 = (ref32.double() - ref64).abs().max().item()
    assert e_tc <= 4 * e_32 + 2e-6 * peak, (e_tc, e_32, peak)
    rel = ((out.double() - ref64).norm() / ref64.norm()).item()
    assert rel <= 5e-6, rel


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (128 * 150 + 128, 256, 256), (384, 256, 44), (256, 512, 256), (256, 48, 256),
                                   (256, 260, 388)])
@pytest.mark.parametrize("epi", [0, 1])
def test_forward(eng, M, N, K, epi):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda") * 0.1
    out = _gemm(eng, 0, epi, M, N, K, A, W, torch.full((M, N), float("nan"), device="cuda"), b)
    ref64 = A.double() @ W.double().t() + b.double()
    ref32 = A @ W.t() + b
    if epi == 1:
        ref64, ref32 = torch.relu(ref64), torch.relu(ref32)
    _check(out, ref64, ref32)


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (128 * 149, 256, 256), (256, 44, 256), (256, 256, 512), (384, 388, 260)])
@pytest.mark.parametrize("epi", [2, 3])
def test_dx(eng, M, N, K, epi):
    """C[M,N] = dY[M,K] W[K,N] (W stored [out = K, in = N]), optionally masked by the layer input"""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(M + N)
    dY = torch.randn(M, K, device="cuda")
    W = torch.randn(K, N, device="cuda") / K ** 0.5
    X = torch.relu(torch.randn(M, N, device="cuda"))
    out = _gemm(eng, 1, epi, M, N, K, dY, W, torch.full((M, N), float("nan"), device="cuda"), X if epi == 3 else None)
    ref64 = dY.double() @ W.double()
    ref32 = dY @ W
    if epi == 3:
        ref64, ref32 = ref64 * (X > 0), ref32 * (X > 0)
    _check(out, ref64, ref32)


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (128 * 300, 256, 256), (1024, 256, 44), (2048, 512, 512), (512, 132, 260)])
def test_dw(eng, M, N, K):
    """C[N,K] += dY[M,N]^T X[M,K]"""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(M + K)
    dY = torch.randn(M, N, device="cuda")
    X = torch.relu(torch.randn(M, K, device="cuda"))
    base = torch.randn(N, K, device="cuda")
    db = torch.ones(N, device="cuda")
    out = _gemm(eng, 2, 2, M, N, K, dY, X, base.clone(), db)
    ref64 = base.double() + dY.double().t() @ X.double()
    ref32 = base + dY.t() @ X
    _check(out, ref64, ref32)
    # the kernel also accumulates the column sums of dY (the bias gradient)
    ref_db = 1.0 + dY.double().sum(0)
    assert (db.double() - ref_db).abs().max().item() <= 1e-5 * (dY.abs().sum(0).max().item() + 1.0)
