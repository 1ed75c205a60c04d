"""GPU tensor2img / PSNR against the oracle restatement of basicsr/utils/img_util.py:59-121 and
basicsr/metrics/psnr_ssim.py:9-61: bit-exact uint8 images, bit-identical PSNR values."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,crop", [((2, 3, 3, 40, 56), 0), ((1, 2, 3, 33, 47), 4), ((1, 1, 3, 720, 1280), 0)])
def test_psnr_frames_bit_identical_to_oracle(shape, crop):
    from oracle import refid_oracle as O
    from refid_b200 import metrics
    g = torch.Generator().manual_seed(21)
    gt = torch.rand(shape, generator=g)
    pred = gt + 0.05 * torch.randn(shape, generator=g)  # values outside [0,1] exercise the clamp
    pred.view(-1)[::5] = (torch.arange(pred.numel() // 5 + 1)[: pred.view(-1)[::5].numel()] % 256 + 0.5) / 255.0  # rounding ties
    mine = metrics.psnr_frames(pred.cuda(), gt.cuda(), crop_border=crop)
    P, G = pred.reshape(-1, *shape[-3:]), gt.reshape(-1, *shape[-3:])
    assert len(mine) == P.shape[0]
    for f in range(P.shape[0]):
        a, b = O.tensor2img_uint8(P[f]), O.tensor2img_uint8(G[f])
        assert mine[f] == O.psnr_uint8(a, b, crop), (f, mine[f], O.psnr_uint8(a, b, crop))
    assert metrics.psnr_frames(gt.cuda(), gt.cuda()) == [float("inf")] * P.shape[0]


def test_tensor2img_matches_oracle_layout():
    from oracle import refid_oracle as O
    from refid_b200 import metrics
    t = torch.rand(3, 37, 52, generator=torch.Generator().manual_seed(2)) * 1.2 - 0.1
    img = metrics.tensor2img(t.cuda())
    ref = O.tensor2img_uint8(t).permute(1, 2, 0).numpy()[:, :, ::-1]  # HWC, RGB -> BGR (img_util.py:105-111)
    assert img.dtype == np.uint8 and img.shape == (37, 52, 3) and np.array_equal(img, ref)
    assert np.array_equal(metrics.tensor2img(t.cuda()[None], rgb2bgr=False), ref[:, :, ::-1])
    gray = metrics.tensor2img(t.cuda()[:1])
    assert gray.shape == (37, 52) and np.array_equal(gray, ref[:, :, 2])
    with pytest.raises(RuntimeError):
        metrics.tensor2img(t)
    with pytest.raises(NotImplementedError):
        metrics.tensor2img(t.cuda(), out_type=np.float32)


def test_gpu_metric_matches_reference_golden_vectors():
    """The GPU kernel against vectors of the UNMODIFIED reference `tensor2img` / `calculate_psnr`
    (tests/golden/make_psnr_golden.py): uint8 images bit-identical, PSNR to 1e-12 (crop border, ties, inf, max_value = 1)."""
    import os
    from refid_b200 import metrics
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "psnr_cases.npz"))
    for n in sorted({k.split(".")[0] for k in z.files}):
        p, q = torch.from_numpy(z[n + ".pred"]).cuda(), torch.from_numpy(z[n + ".gt"]).cuda()
        assert np.array_equal(metrics.tensor2img(p), z[n + ".img_pred"]), n
        assert np.array_equal(metrics.tensor2img(q), z[n + ".img_gt"]), n
        want, got = float(z[n + ".psnr"]), metrics.psnr_frames(p, q, crop_border=int(z[n + ".crop"]))[0]
        assert (np.isinf(want) and np.isinf(got)) or abs(want - got) < 1e-12, (n, want, got)
