"""Golden vectors of the reference's DATA PATH: `inputs.prepare_images` (inputs.py:107-127) run UNMODIFIED (CPU) on the first five
views of its default dataset data/planar/cat_batch3 — LANCZOS thumbnail to the patch size, mask inversion `(im < 0.5)`, 5x5
erosion, grey images, Sobel/Gauss edge labels (OpenCV, float64) and the kornia-normalised ground-truth homographies with the
reference's swapped (width, height) arguments (inputs.py:104).

TEST INFRASTRUCTURE ONLY.  Run in the build container:  `python oracle/gen_golden_data.py`
The input files (5 views, 5 masks, gt.png, H_0_1..4.mat: data, not code) are copied to tests/golden/cat_batch3/ so that the tests
feed the same bytes through marf_b200/inputs.py on any box; this script writes tests/golden/data_cat_batch3.npz.
kornia is absent here: `normalize_homography` is restated in oracle/ref_runner.py from kornia's documented formula, so the
`gt_hom` vector pins the restatement, not the library."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fixtures as fx            # noqa: E402
import ref_runner                # noqa: E402

ROOT = os.path.dirname(HERE)
DATA = os.path.join(ROOT, "tests", "golden", "cat_batch3")
OUT = os.path.join(ROOT, "tests", "golden", "data_cat_batch3.npz")


def main():
    planar, warp_mod, inputs_mod, edict = ref_runner.load_reference()
    opt = ref_runner.make_opt(edict, use_edges=True, use_homographies=True, dataset="cat_batch3")
    B = opt.batch_size
    images = inputs_mod.prepare_images(
        opt,
        fps_images=[f"{DATA}/{i}.png" for i in range(B)],
        fps_masks=[f"{DATA}/{i}-m.png" for i in range(B)],
        fp_gt=f"{DATA}/gt.png",
        fps_hom=[f"{DATA}/H_0_{i}.mat" for i in range(1, B)],
        edges=True)
    flat = {}

    def u8(t):          # to_tensor of an 8-bit image: exactly k/255 in float32
        a = t.numpy()
        q = np.round(a * 255).astype(np.uint8)
        assert np.array_equal(q.astype(np.float32) / 255, a.astype(np.float32)), "not an 8-bit image"
        return q
    flat["rgb_u8"] = u8(images.rgb)
    flat["gray_u8"] = u8(images.gray)
    flat["gt_u8"] = u8(images.gt)
    flat["masks_bits"] = np.packbits(images.masks.numpy().astype(np.uint8))
    flat["masks_shape"] = np.array(images.masks.shape)
    flat["masks_eroded_bits"] = np.packbits(images.masks_eroded.numpy().astype(np.uint8))
    assert set(np.unique(images.masks.numpy())) <= {0.0, 1.0} and set(np.unique(images.masks_eroded.numpy())) <= {0.0, 1.0}
    flat["gt_hom"] = images.gt_hom.numpy()
    assert images.edges.dtype == torch.float64
    for k, v in fx.digest(images.edges, n_probe=4096).items():
        flat[f"edges.{k}"] = np.asarray(v)
    flat["edges_shape"] = np.array(images.edges.shape)
    np.savez_compressed(OUT, **flat)
    print(f"wrote {OUT} ({os.path.getsize(OUT)/1024:.1f} KiB); rgb {tuple(images.rgb.shape)} masks {tuple(images.masks.shape)} "
          f"edges {tuple(images.edges.shape)} {images.edges.dtype} gt {tuple(images.gt.shape)} gt_hom {tuple(images.gt_hom.shape)}")


if __name__ == "__main__":
    main()
