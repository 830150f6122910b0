"""Generate tests/golden/preproc_pil.npz by running the REAL PIL / torchvision pipeline the reference
uses for its network inputs (api/sampler/snippet_sampler.py:156-185, api/utils/data_utils.py:71-120,
api/utils/model_utils.py:26-40) on seeded synthetic 112x112 face crops.  Run in the dev container:

    python oracle/make_golden_preproc.py

Also asserts that oracle/pil_preproc.py reproduces PIL bit for bit before writing.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pil_preproc as P  # noqa: E402


def make_crops():
    rng = np.random.default_rng(7)
    crops = rng.integers(0, 256, (6, 112, 112, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:112, 0:112]
    crops[1] = np.stack([(127 + 120 * np.sin(xx / 7.0 + c) * np.cos(yy / 11.0)).astype(np.uint8) for c in range(3)], -1)
    crops[2] = 255
    crops[3] = 0
    crops[4] = ((xx + yy) % 2 * 255)[..., None].astype(np.uint8)          # checkerboard: ringing clips at 0/255
    crops[5, 30:80, 20:90] = (200, 40, 90)                                 # flat patch with sharp edges in noise
    return crops


def pil_pipeline(crops):
    from PIL import Image
    import torchvision.transforms as T
    to_rgb = T.Compose([T.Resize(256), T.CenterCrop((224, 224))])
    full = T.Compose([T.Resize(256), T.CenterCrop((224, 224)), T.ToTensor(), lambda x: x * 255.0,
                      T.Normalize(mean=list(P.RESNET_MEAN), std=[1, 1, 1])])
    scale = T.Resize(48, Image.LANCZOS)                                    # GroupScale(48), data_utils.py:80-81
    gray, rgb_u8, rgb_f = [], [], []
    for c in crops:
        im = Image.fromarray(c, 'RGB')
        g = np.stack([scale(im.convert('L'))], axis=2)                     # Stack, data_utils.py:94-95
        gray.append(torch.from_numpy(g).permute(2, 0, 1).contiguous().float().div(255)[0].numpy())
        rgb_u8.append(np.moveaxis(np.asarray(to_rgb(im)), -1, 0))
        rgb_f.append(full(im).numpy())
    return np.stack(gray), np.stack(rgb_u8), np.stack(rgb_f)


def main():
    crops = make_crops()
    gray, rgb_u8, rgb_f = pil_pipeline(crops)
    assert np.array_equal(P.crops_to_gray(crops), gray), "oracle gray != PIL"
    assert np.array_equal(P.crops_to_rgb(crops), rgb_f), "oracle rgb != torchvision"
    out = os.path.join(ROOT, "tests", "golden", "preproc_pil.npz")
    # the full fp32 RGB tensor is 3.6 MB of noise; keep PIL's uint8 image plus fp32 rows 0..7 (every
    # possible uint8 value x channel occurs there, pinning the u8 -> float arithmetic)
    np.savez_compressed(out, crops=crops, gray=gray, rgb_u8=rgb_u8, rgb_f32_rows=rgb_f[:, :, :8, :])
    import PIL
    import torchvision
    print("wrote", out, os.path.getsize(out), "bytes; PIL", PIL.__version__, "torchvision", torchvision.__version__)


if __name__ == "__main__":
    main()
