"""The drop-in command line on the reference's own data: `python train.py --model=planar --yaml=planar ...` (reference README /
train.py:11-31) in a scratch directory that holds what the reference's working directory holds — options/, model/, train.py and
data/planar/cat_batch3/ (the fixture copy of the first five views) — for a short run, in the reference's default configuration
(disk masks + edge term, fp32) and in the tensor-core mode with the learned mask."""
import os
import re
import subprocess
import sys

import pytest
import torch

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _workdir(tmp_path):
    for name in ("options", "model", "marf_b200", "train.py", "include"):
        os.symlink(os.path.join(ROOT, name), tmp_path / name)
    (tmp_path / "data" / "planar").mkdir(parents=True)
    os.symlink(os.path.join(cases.GOLDEN, "cat_batch3"), tmp_path / "data" / "planar" / "cat_batch3")
    return tmp_path


@pytest.mark.parametrize("flags", [
    ["--barf_c2f=[0,0.4]"],                                                    # the README command (fp32, disk masks, edges, H_0_i.mat)
    ["--use_implicit_mask", "--precision=bf16", "--use_homographies!", "--fused_optimizer"],
])
def test_train_cli_runs_on_reference_data(tmp_path, flags):
    wd = _workdir(tmp_path)
    cmd = [sys.executable, "train.py", "--model=planar", "--yaml=planar", "--name=cli", "--seed=3", "--max_iter=120",
           "--freq.vis=60", "--freq.scalar=20", "--tb!"] + flags
    r = subprocess.run(cmd, cwd=wd, capture_output=True, text=True, timeout=600, stdin=subprocess.DEVNULL)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert "TRAINING DONE" in r.stdout
    out = [d for d, _, files in os.walk(wd / "output") if "options.yaml" in files]
    assert len(out) == 1
    frames = sorted(os.listdir(os.path.join(out[0], "vis")))
    assert frames == ["0.png", "1.png", "2.png"], frames                      # step 0 + every freq.vis iterations
    ck = torch.load(os.path.join(out[0], "model.ckpt"), weights_only=True)
    assert ck["it"] == 120 and torch.isfinite(ck["graph"]["warp_param.weight"]).all()
    assert float(ck["graph"]["warp_param.weight"][0].abs().max()) == 0.0      # warp.fix_first
    assert float(ck["graph"]["warp_param.weight"][1:].abs().max()) > 1e-3     # the other patches moved
    assert float(ck["graph"]["neural_image.progress"]) == pytest.approx(1.0)
