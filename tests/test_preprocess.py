"""Image preprocessing: the oracle's restatement of Pillow's resampler vs the reference's own torchvision pipeline
(CPU), and the CUDA kernels vs that pipeline (GPU, bit-exact)."""
import numpy as np
import pytest
import torch

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
CASES = [(500, 700, 3), (2544, 3056, 1), (1000, 900, 1), (300, 200, 3), (384, 384, 3), (385, 1200, 1), (3000, 2000, 3),
         (391, 389, 1), (4000, 384, 1)]


def _img(h, w, c, seed):
    rng = np.random.default_rng(seed)
    smooth = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2) + ((3,) if c == 3 else ()), dtype=np.uint8)
    img = np.kron(smooth, np.ones((8, 8) + ((1,) if c == 3 else ()), dtype=np.uint8))[:h, :w]   # blocky + noise
    noise = rng.integers(-20, 21, img.shape)
    return np.clip(img.astype(np.int64) + noise, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("h,w,c", CASES)
def test_restated_resampler_equals_torchvision_pipeline(h, w, c):
    from oracle import preprocess as P
    img = _img(h, w, c, seed=h * 7 + w)
    a = P.reference_test_transforms(img, MEAN, STD)
    b = P.test_transforms_restated(img, MEAN, STD)
    assert a.shape == (3, 384, 384)
    assert np.array_equal(a, b)


@pytest.mark.gpu
def test_gpu_preprocess_bit_exact_vs_reference_pipeline():
    from cxrmate_b200.engine import Engine
    from cxrmate_b200.preprocess import TestTransforms
    from oracle import preprocess as P
    e = Engine(dtype="fp32", image_size=64, max_studies=2, max_images=2, max_prompt=8, max_new_tokens=4, rwd_layers=0, enc_chunk=2,
               dec_layers=1, cvt_depth=(1, 1, 1))
    try:
        tf = TestTransforms(e, MEAN, STD)
        for i, (h, w, c) in enumerate(CASES):
            img = _img(h, w, c, seed=h * 7 + w)
            ref = P.reference_test_transforms(img, MEAN, STD)
            src = torch.from_numpy(img)
            got = tf(src.cuda() if i % 2 else src)            # device-resident and host inputs
            torch.cuda.synchronize()
            assert np.array_equal(got.cpu().numpy(), ref), (h, w, c, np.abs(got.cpu().numpy() - ref).max())
        # a padded study batch: missing slots stay exactly zero (how the reference marks padding)
        px = tf.batch([[_img(500, 700, 3, 1), _img(600, 500, 1, 2)], [_img(400, 400, 1, 3)]], max_images=3)
        assert px.shape == (2, 3, 3, 384, 384)
        assert px[0, 2].abs().max().item() == 0.0 and px[1, 1:].abs().max().item() == 0.0 and px[1, 0, 0, 0, 0].item() != 0.0
    finally:
        e.close()
