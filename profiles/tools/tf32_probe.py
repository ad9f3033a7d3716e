"""Accuracy of the 3xTF32 GEMMs vs torch fp32 / float64 (prints; run on the GPU box)."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import planar_oracle as po
import gpu_util
from marf_b200 import _lib as L

torch.backends.cuda.matmul.allow_tf32 = False
lib = L.load()
eng = gpu_util.make_engine(po.PlanarConfig(H=40, W=56, patch_H=20, patch_W=28, batch_size=2, use_masks=False), "fp32")
st = torch.cuda.current_stream().cuda_stream
for M, N, K in [(1 << 16, 256, 256), (1 << 16, 256, 64), (1 << 16, 256, 1024)]:
    torch.manual_seed(0)
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
    out = torch.empty(M, N, device="cuda")
    rc = lib.marf_tf32_gemm(eng.handle, 0, 2, M, N, K, A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), N, None, 0, C.c_void_p(st))
    assert rc == 0
    r64 = A.double() @ W.double().t()
    r32 = A @ W.t()
    torch.backends.cuda.matmul.allow_tf32 = True
    rtf = A @ W.t()
    torch.backends.cuda.matmul.allow_tf32 = False
    f = lambda x: ((x.double() - r64).norm() / r64.norm()).item()
    g = lambda x: ((x.double() - r64).abs().max() / r64.abs().max()).item()
    print(f"M={M} N={N} K={K}: rel-L2 err  3xTF32 {f(out):.2e}  torch fp32 {f(r32):.2e}  torch tf32 {f(rtf):.2e}   max/peak  {g(out):.2e} {g(r32):.2e} {g(rtf):.2e}")
    # bias of the error (truncating accumulation shows as a non-zero mean)
    print(f"    mean signed err / rms: 3xTF32 {((out.double()-r64).mean()/r64.pow(2).mean().sqrt()).item():.2e}  fp32 {((r32.double()-r64).mean()/r64.pow(2).mean().sqrt()).item():.2e}")
eng.close()

# ---- timing of the three GEMM forms (CUDA events around the diagnostic entry; it synchronises after the kernel)
eng = gpu_util.make_engine(po.PlanarConfig(H=40, W=56, patch_H=20, patch_W=28, batch_size=2, use_masks=False), "fp32")
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return min(ts)
for M in (216064, 1 << 19):
    N = K = 256
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / 16; X = torch.relu(torch.randn(M, N, device="cuda"))
    out = torch.empty(M, N, device="cuda"); b = torch.zeros(N, device="cuda"); dW = torch.zeros(N, K, device="cuda")
    call = lambda mode, epi, Cout, aux, ldaux: lib.marf_tf32_gemm(eng.handle, mode, epi, M, N, K, A.data_ptr(), K, (W if mode < 2 else X).data_ptr(), K,
                                                                 Cout.data_ptr(), Cout.stride(0), aux.data_ptr() if aux is not None else None, ldaux, C.c_void_p(st))
    fl = 2.0 * M * N * K
    for name, fn in [("fwd bias+relu", lambda: call(0, 1, out, b, 0)), ("dX plain (incl. W transpose)", lambda: call(1, 2, out, None, 0)),
                     ("dX relu-mask (incl. W transpose)", lambda: call(1, 3, out, X, N)), ("dW + db", lambda: call(2, 2, dW, b, 0))]:
        us = timed(fn)
        print(f"M={M}: {name:34s} {us:8.1f} us  {fl / us * 1e-6:7.1f} TFLOP/s (fp32-equivalent)")
eng.close()
