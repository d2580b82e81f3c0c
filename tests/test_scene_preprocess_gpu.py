"""load_images on the GPU: d3r_image_resize_crop_normalize through the C ABI vs the host pipeline of
load_images (Pillow resize / crop + ImgNorm, itself equal to the unmodified reference: tests/test_image_preprocess.py), the
CPU oracle and the reference's golden outputs -- bit-exact (integer / byte work) -- and inference() on images that were born
on the device."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from dust3r_b200.utils import image as img_mod
from dust3r_b200.utils.synth import synth_photo

pytestmark = pytest.mark.gpu

CASES = [  # (H, W, size, square_ok)
    (150, 200, 128, False), (200, 150, 128, False), (130, 130, 128, False), (130, 130, 128, True), (37, 53, 224, False),
    (300, 170, 224, False), (480, 640, 512, False), (640, 480, 512, False), (384, 512, 512, False), (97, 1003, 512, False),
    (601, 397, 224, False), (3000, 4000, 512, False), (2000, 1500, 224, False),
]


def _write_png(tmp_path, arr, name):
    import PIL.Image
    path = os.path.join(str(tmp_path), name)
    PIL.Image.fromarray(arr).save(path)
    return path


@pytest.mark.timeout(600)
def test_gpu_load_images_equals_host_load_images(cuda_device, tmp_path):
    paths = [_write_png(tmp_path, synth_photo(h, w, seed=30 + k), f'{k:02d}.png') for k, (h, w, _, _) in enumerate(CASES)]
    for k, (h, w, size, sq) in enumerate(CASES):
        host = img_mod.load_images([paths[k]], size=size, square_ok=sq, verbose=False)[0]
        gpu = img_mod.load_images([paths[k]], size=size, square_ok=sq, verbose=False, device=cuda_device)[0]
        assert gpu['img'].is_cuda and gpu['img'].dtype == torch.float32
        assert tuple(gpu['img'].shape) == tuple(host['img'].shape), (h, w, size, sq)
        assert torch.equal(gpu['img'].cpu(), host['img']), (h, w, size, sq, int((gpu['img'].cpu() != host['img']).sum()))
        assert np.array_equal(gpu['true_shape'], host['true_shape']) and gpu['idx'] == host['idx'] == 0
    # a folder: same ordering, idx and instance as the host path
    host = img_mod.load_images(str(tmp_path), size=512, verbose=False)
    gpu = img_mod.load_images(str(tmp_path), size=512, verbose=False, device=cuda_device)
    assert [v['instance'] for v in gpu] == [v['instance'] for v in host] and len(gpu) == len(CASES)
    for a, b in zip(gpu, host):
        assert torch.equal(a['img'].cpu(), b['img'])


@pytest.mark.timeout(600)
def test_gpu_preprocess_equals_oracle_and_reference_golden(cuda_device):
    from oracle import image_oracle as io
    gold = np.load(os.path.join(GOLDEN, 'load_images.npz'))
    lut = img_mod.norm_lut().numpy()
    for k in range(len([f for f in gold.files if f.endswith('|in')])):
        size, square_ok = (int(v) for v in gold[f'{k}|args'])
        got = img_mod.preprocess_image_u8(gold[f'{k}|in'], size, bool(square_ok), cuda_device).cpu().numpy()
        assert np.array_equal(got, np.moveaxis(lut[gold[f'{k}|out_u8']], -1, 0)[None]), k
    rng = np.random.default_rng(3)
    for h, w, size in ((211, 317, 512), (900, 700, 224), (64, 64, 96)):
        photo = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)        # white noise: heavy clipping of the filter overshoot
        want, true_shape = io.load_image_oracle(photo, size)
        got = img_mod.preprocess_image_u8(torch.from_numpy(photo).to(cuda_device), size, device=cuda_device)   # device-resident source
        assert np.array_equal(got.cpu().numpy(), want) and tuple(got.shape[-2:]) == tuple(true_shape[0])


@pytest.mark.timeout(900)
def test_inference_on_device_resident_images_equals_host_images(cuda_device, tmp_path):
    """Images resized on the GPU are used in place by inference() (no round trip through the host) and give the bits host
    images give."""
    from test_forward_gpu import _build, _small_cfgs
    from dust3r_b200.image_pairs import make_pairs
    from dust3r_b200.inference import inference
    cfg, H, W = _small_cfgs()['small_dpt']
    net, _ = _build(cfg, 11, cuda_device)
    for k in range(3):
        _write_png(tmp_path, synth_photo(3 * H, 3 * W, seed=40 + k), f'{k}.png')      # -> exactly H x W after the resize
    size = max(H, W)
    host = img_mod.load_images(str(tmp_path), size=size, verbose=False)
    gpu = img_mod.load_images(str(tmp_path), size=size, verbose=False, device=cuda_device)
    for a, b in zip(gpu, host):
        assert tuple(a['img'].shape) == (1, 3, H, W) and torch.equal(a['img'].cpu(), b['img'])
    for sym in (True, False):
        a = inference(make_pairs(host, symmetrize=sym), net, cuda_device, batch_size=4, verbose=False)
        b = inference(make_pairs(gpu, symmetrize=sym), net, cuda_device, batch_size=4, verbose=False)
        # private device copies of every image: the non-deduplicated upload path
        private = [(dict(x, img=x['img'].clone()), dict(y, img=y['img'].clone())) for x, y in make_pairs(gpu, symmetrize=sym)]
        c = inference(private, net, cuda_device, batch_size=4, verbose=False)
        for other in (b, c):
            assert a['view1']['idx'] == other['view1']['idx'] and a['view2']['idx'] == other['view2']['idx']
            assert not other['view1']['img'].is_cuda and torch.equal(a['view1']['img'], other['view1']['img'])
            for which, key in (('pred1', 'pts3d'), ('pred1', 'conf'), ('pred2', 'pts3d_in_other_view'), ('pred2', 'conf')):
                assert torch.equal(a[which][key], other[which][key]), (sym, which, key)
