import ctypes as C, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from marf_b200 import _lib as L
lib = L.load()
rows, K, N = 128*148*12, 256, 256
A = torch.randn(rows, K, device='cuda'); W = torch.randn(N, K, device='cuda')/16; b = torch.randn(N, device='cuda')*0.1
out = torch.zeros(rows, N, device='cuda')
torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    rc = lib.marf_tc_selftest(0, 0, rows, K, N, A.data_ptr(), W.data_ptr(), b.data_ptr(), out.data_ptr(), C.c_void_p(st))
print('rc', rc)
