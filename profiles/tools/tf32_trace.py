"""clock64 trace of CTA 0 of one 3xTF32 GEMM (MARF_T32_TRACE=1): python profiles/tools/tf32_trace.py <mode> <epi>"""
import ctypes as C
import os, sys
if len(sys.argv) < 4:
    os.environ["MARF_T32_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import planar_oracle as po
import gpu_util
from marf_b200 import _lib as L
lib = L.load()
eng = gpu_util.make_engine(po.PlanarConfig(H=40, W=56, patch_H=20, patch_W=28, batch_size=2, use_masks=False), "fp32")
st = torch.cuda.current_stream().cuda_stream
mode, epi = int(sys.argv[1]), int(sys.argv[2])
M, N, K = 216064, 256, 256
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / 16; X = torch.relu(torch.randn(M, N, device="cuda"))
out = torch.zeros(M if mode < 2 else N, N, device="cuda"); b = torch.zeros(N, device="cuda")
aux = X if epi == 3 else b
for _ in range(2):
    rc = lib.marf_tf32_gemm(eng.handle, mode, epi, M, N, K, A.data_ptr(), K, (W if mode < 2 else X).data_ptr(), K, out.data_ptr(), N,
                            aux.data_ptr(), N if epi == 3 else 0, C.c_void_p(st))
    assert rc == 0
eng.close()
