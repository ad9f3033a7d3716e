"""The real-data path: marf_b200/inputs.py against golden vectors the UNMODIFIED reference produced from the same files
(inputs.prepare_images, inputs.py:107-127, on the first five views of data/planar/cat_batch3; oracle/gen_golden_data.py).
The input PNG / .mat files are fixtures under tests/golden/cat_batch3/ (data, not code)."""
import os

import numpy as np
import pytest
import torch

import cases
import fixtures as fx

DATA = os.path.join(cases.GOLDEN, "cat_batch3")


def _opt(device):
    from marf_b200.attrdict import AttrDict
    return AttrDict(H=360, W=480, patch_H=180, patch_W=240, batch_size=5, use_cropped_images=True, device=device)


def _prepare(device, edges):
    from marf_b200 import inputs
    B = 5
    return inputs.prepare_images(_opt(device), fps_images=[f"{DATA}/{i}.png" for i in range(B)],
                                 fps_masks=[f"{DATA}/{i}-m.png" for i in range(B)], fp_gt=f"{DATA}/gt.png",
                                 fps_hom=[f"{DATA}/H_0_{i}.mat" for i in range(1, B)], edges=edges)


def _check_static(im, g):
    """Everything but the edge labels: bit-exact (8-bit images, binary masks), homographies to fp32 rounding."""
    assert np.array_equal(im.rgb.cpu().numpy(), g["rgb_u8"].astype(np.float32) / 255)           # LANCZOS thumbnail, inputs.py:26-29
    assert np.array_equal(im.gray.cpu().numpy(), g["gray_u8"].astype(np.float32) / 255)
    assert np.array_equal(im.gt.cpu().numpy(), g["gt_u8"].astype(np.float32) / 255)
    shape = tuple(g["masks_shape"])
    n = int(np.prod(shape))
    masks = np.unpackbits(g["masks_bits"])[:n].reshape(shape).astype(np.float32)
    eroded = np.unpackbits(g["masks_eroded_bits"])[:n].reshape(shape).astype(np.float32)
    assert np.array_equal(im.masks.cpu().numpy(), masks)                                          # (im < 0.5) inversion, inputs.py:30-31
    assert 0.5 < masks.mean() < 1.0                                                                # white = occluded on disk, 1 = valid here
    assert np.array_equal(im.masks_eroded.cpu().numpy(), eroded)                                  # cv2.erode 5x5, inputs.py:71-85
    # kornia normalisation with the reference's swapped (width, height) arguments (inputs.py:104)
    np.testing.assert_allclose(im.gt_hom.cpu().numpy(), g["gt_hom"], rtol=2e-6, atol=2e-6)
    assert np.array_equal(im.gt_hom[0].cpu().numpy(), np.eye(3, dtype=np.float32))


def test_prepare_images_matches_reference_on_cat_batch3():
    g = cases.load_golden("data_cat_batch3")
    im = _prepare("cpu", edges=None)
    assert im.edges is None
    _check_static(im, g)


@pytest.mark.gpu
def test_prepare_images_with_edges_matches_reference_on_cat_batch3():
    """The same with the Sobel/Gauss edge labels (marf_compute_edges on the GPU vs OpenCV in the reference, float64)."""
    g = cases.load_golden("data_cat_batch3")
    im = _prepare("cuda:0", edges=True)
    _check_static(im, g)
    assert im.edges.dtype == torch.float64 and tuple(im.edges.shape) == tuple(g["edges_shape"])
    d = fx.digest(im.edges, n_probe=4096)
    assert d["size"] == int(g["edges.size"])
    np.testing.assert_allclose(d["sample"], g["edges.sample"], rtol=0, atol=1e-12)
    np.testing.assert_allclose([d["norm"], d["sum"], d["proj"]], [float(g["edges.norm"]), float(g["edges.sum"]), float(g["edges.proj"])],
                               rtol=1e-12, atol=1e-9)


@pytest.mark.gpu
def test_model_load_dataset_reads_the_reference_layout(tmp_path, monkeypatch):
    """Model.load_dataset (model/planar.py:59-78) through the plugin: data/planar/<dataset>/{i.png, i-m.png, gt.png, H_0_i.mat}."""
    from marf_b200 import options, planar
    root = tmp_path / "data" / "planar"
    root.mkdir(parents=True)
    os.symlink(DATA, root / "cat_batch3")
    opt = options.load_options(os.path.join(os.path.dirname(cases.GOLDEN), "..", "options", "planar.yaml"))
    opt.update(model="planar", yaml="planar", device="cuda:0", output_path=str(tmp_path / "out"), tb=None, seed=3, world_size=1, rank=0)
    monkeypatch.chdir(tmp_path)
    m = planar.Model(opt)
    m.load_dataset()
    _check_static(m.images, cases.load_golden("data_cat_batch3"))
    assert m.images.edges is not None and tuple(m.images.edges.shape) == (5, 1, 180, 240)
