"""GPU: the measurement hooks bench.py relies on (marf_profile / marf_profile_read) and the debug read-back."""
import ctypes as C

import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


def test_profile_records_every_tensor_core_launch():
    import gpu_util
    cfg, params, images, it, progress, g = cases.build_case("implicit")
    eng = gpu_util.make_engine(cfg, "bf16")
    gpu_util.run_step(eng, cfg, params, images, it, progress)            # warm-up, not recorded
    assert all(n == 0 for _, n in eng.profile_read().values())
    eng.profile(True)
    for _ in range(3):
        gpu_util.run_step(eng, cfg, params, images, it, progress)
    rec = eng.profile_read()
    eng.profile(False)
    for k in ("k_tc_chain<fwd>", "k_tc_bwd", "k_tc_gemm<64,warp_grad>"):      # (k_tc_bwd = the dX chains + every dW GEMM, one launch)
        ms, n = rec[k]
        assert n == 3, (k, n)
        assert 0.0 < ms < 50.0, (k, ms)
    assert rec["k_tc_dw"][1] == 0 and rec["k_tc_chain<dx>"][1] == 0
    assert all(n == 0 for _, n in eng.profile_read().values())           # read clears
    gpu_util.run_step(eng, cfg, params, images, it, progress)
    assert all(n == 0 for _, n in eng.profile_read().values())           # disabled: nothing recorded
    eng.close()


def test_debug_read_returns_the_resident_layer_input():
    """act[1] = relu(X0 W0^T + b0) in bf16: non-negative, about half of it zero, and consistent with the 1-bit ReLU masks
    the dX chain uses (checked indirectly: the step's gradients pass parity in test_gpu_parity_bf16)."""
    import gpu_util
    cfg, params, images, it, progress, g = cases.build_case("mid_mask")
    eng = gpu_util.make_engine(cfg, "bf16")
    gpu_util.run_step(eng, cfg, params, images, it, progress)
    rows = (cfg.batch_size * cfg.h * cfg.w + 127) // 128 * 128
    t = torch.full((rows, 256), -1.0, device="cuda")
    rc = eng.lib.marf_debug_read_bf16(eng.handle, 0, 0, 1, C.c_void_p(t.data_ptr()), rows, None)
    assert rc == 0
    n = cfg.batch_size * cfg.h * cfg.w
    assert (t[:n] >= 0).all()
    frac0 = (t[:n] == 0).float().mean().item()
    assert 0.2 < frac0 < 0.8, frac0
    assert eng.lib.marf_debug_read_bf16(eng.handle, 0, 0, 99, C.c_void_p(t.data_ptr()), rows, None) != 0
    eng.close()
