"""f-2 / a-2 augmentation branch: ColorJitter (given parameters) + Normalize on the device against torchvision's
functional ops (what the reference runs).  Same float32 operation order; the contrast step's image mean is summed in a
different order (double, fixed shape) -> tolerance 2e-6 before / 1e-5 after Normalize."""
import itertools

import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu


def _params(orders, facs, flags=None):
    from lighthand_b200 import _lib
    recs = (_lib.JitterParams * len(orders))()
    for i, (o, f) in enumerate(zip(orders, facs)):
        recs[i].apply = 1 if flags is None else int(flags[i])
        for k in range(4):
            recs[i].order[k] = int(o[k])
            recs[i].factor[k] = float(f[k])
            recs[i].one_minus[k] = 1.0 - float(f[k])
    return recs


def test_every_order_random_factors():
    from lighthand_b200 import ops
    rng = np.random.default_rng(41)
    orders = list(itertools.permutations(range(4)))
    n = len(orders)
    imgs = rng.uniform(0, 1, (n, 3, 40, 56)).astype(np.float32)
    facs = np.stack([rng.uniform(0.5, 1.5, n), rng.uniform(0.5, 1.5, n), rng.uniform(0.5, 1.5, n), rng.uniform(-0.5, 0.5, n)], 1)
    facs = facs.astype(np.float32).astype(np.float64)      # the host draw is a float32 value (torch.empty(1).uniform_)
    x = torch.from_numpy(imgs).cuda()
    raw = ops.color_jitter_normalize(x, _params(orders, facs), mean=None).cpu().numpy()
    nrm = ops.color_jitter_normalize(x, _params(orders, facs)).cpu().numpy()
    for i in range(n):
        want = R.color_jitter_sample(imgs[i], orders[i], *facs[i], mean=None)
        np.testing.assert_allclose(raw[i], want, rtol=0, atol=2e-6, err_msg=f"order {orders[i]} factors {facs[i]}")
        np.testing.assert_allclose(nrm[i], R.color_jitter_sample(imgs[i], orders[i], *facs[i]), rtol=0, atol=1e-5)


def test_edge_pixels_and_flags():
    """Grey pixels (max == min), pure black / white, saturated primaries, hue wrap at +-0.5, factors at the ends of
    their ranges; samples whose flag is off are only normalised; in-place operation."""
    from lighthand_b200 import ops
    rng = np.random.default_rng(42)
    n = 6
    imgs = rng.uniform(0, 1, (n, 3, 16, 16)).astype(np.float32)
    imgs[:, :, 0, :] = imgs[:, :1, 0, :]                     # grey row
    imgs[:, :, 1, :4] = 0.0
    imgs[:, :, 1, 4:8] = 1.0
    imgs[:, 0, 2, :] = 1.0; imgs[:, 1, 2, :] = 0.0; imgs[:, 2, 2, :] = 0.0          # pure red
    imgs[:, 0, 3, :] = imgs[:, 1, 3, :]                       # r == g ties in the max
    orders = [(3, 0, 1, 2), (1, 3, 2, 0), (0, 1, 2, 3), (2, 3, 1, 0), (3, 2, 1, 0), (0, 2, 3, 1)]
    facs = np.array([[0.5, 0.5, 0.5, -0.5], [1.5, 1.5, 1.5, 0.5], [1.0, 1.0, 1.0, 0.0], [0.5, 1.5, 0.5, 0.25],
                     [1.5, 0.5, 1.5, -0.25], [1.25, 0.75, 1.0, 0.125]])
    flags = [1, 1, 1, 0, 1, 1]
    x = torch.from_numpy(imgs).cuda()
    got = ops.color_jitter_normalize(x, _params(orders, facs, flags), out=x)
    assert got.data_ptr() == x.data_ptr()
    got = got.cpu().numpy()
    for i in range(n):
        if flags[i]:
            want = R.color_jitter_sample(imgs[i], orders[i], *facs[i])
        else:
            want = R.color_jitter_sample(imgs[i], (), None, None, None, None)
        np.testing.assert_allclose(got[i], want, rtol=0, atol=1e-5, err_msg=f"sample {i}")


def test_gpu_collate_with_jitter_matches_the_reference_compose():
    """GpuCollate on (image, joints, augment) samples == the reference's two Compose pipelines under the same seed."""
    from torchvision import transforms
    from lighthand_b200.tools.dataset import GpuCollate
    rng = np.random.default_rng(43)
    imgs = rng.integers(0, 256, (5, 56, 56, 3), dtype=np.uint8)
    joints = rng.uniform(0, 56, (5, 21, 2)).astype(np.float32)
    flags = [True, False, True, True, False]
    norm = transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    aug = transforms.Compose([transforms.ToTensor(), transforms.Resize((64, 64), antialias=False),
                              transforms.ColorJitter(brightness=0.5, contrast=0.5, saturation=0.5, hue=0.5), norm])
    plain = transforms.Compose([transforms.ToTensor(), transforms.Resize((64, 64), antialias=False), norm])
    torch.manual_seed(5)
    want = torch.stack([(aug if f else plain)(im) for im, f in zip(imgs, flags)])
    torch.manual_seed(5)
    got, j, hm = GpuCollate(image_size=64, heatmap_size=16)([(imgs[i], joints[i], flags[i]) for i in range(5)])
    torch.testing.assert_close(got.cpu(), want, rtol=0, atol=2e-5)
    assert hm.shape == (5, 21, 16, 16) and j.shape == (5, 21, 2)
