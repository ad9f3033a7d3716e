"""End-point parity, short version (SURVEY §8 row g): 400 iterations of the training loop (model/planar.py:136-209) at the MID
size from a small misalignment (noise_h 0.03, noise_t 0.05: the run registers to ~0.2 px within the 400 iterations), this library
in both precisions against the oracle port on the same GPU, same scene, same init.
The full 3000-iteration run is profiles/r02_endpoint_parity.{json,txt} (tests/endpoint_parity.py).

Tolerances: Adam amplifies last-bit differences, and the weight gradients are summed with reductions whose order changes from run
to run.  Observed over three runs: oracle fp32 0.177 px / 34.54 dB, oracle fp64 0.210 / 34.12, this library fp32 0.226-0.236 px /
34.23-34.26 dB, bf16 0.158-0.207 px / 33.96-34.41 dB — i.e. deviations from the oracle of <= 0.06 px and <= 0.6 dB, the size of the
oracle's own fp32-vs-fp64 difference (0.03 px / 0.4 dB).  Bounds: 0.12 px (BASELINE's criterion is 0.1 px at max_iter) and 1.0 dB."""
import argparse

import pytest

pytestmark = pytest.mark.gpu


def test_endpoint_short_run_matches_oracle():
    import endpoint_parity as ep
    args = argparse.Namespace(batch=3, noise_h=0.03, noise_t=0.05, scene_seed=0, torch_adam=False)
    res = ep.run_scene("c1", 400, 72, 96, ["repo_fp32", "repo_bf16", "oracle_fp32", "oracle_fp64"], args, every=100)
    ref = res["oracle_fp32"]
    assert ref["hist"][0]["corner_px"] > 1.0 and ref["corner_px"] < 0.4          # the run registers
    for arm in ("repo_fp32", "repo_bf16", "oracle_fp64"):
        d = res[arm]["delta_vs_oracle_fp32"]
        print(arm, res[arm]["corner_px"], res[arm]["psnr_last200"], d)
        assert res[arm]["corner_px"] < 0.4, (arm, res[arm]["corner_px"])
        assert d["corner_px"] <= 0.12, (arm, d)
        assert d["psnr_last200"] <= 1.0, (arm, d)
    # the first iteration is a single step from identical parameters: losses agree to fp32 / bf16 step precision
    # (fp32 mode = 3xTF32 tensor-core GEMMs: 2e-6 relative per GEMM; the step tests hold the loss to 2e-5, here 5e-5 after one update)
    assert abs(res["repo_fp32"]["hist"][0]["loss"] - ref["hist"][0]["loss"]) <= 5e-5 * abs(ref["hist"][0]["loss"])
    assert abs(res["repo_bf16"]["hist"][0]["loss"] - ref["hist"][0]["loss"]) <= 2e-2 * abs(ref["hist"][0]["loss"])
