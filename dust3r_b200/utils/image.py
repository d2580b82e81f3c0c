"""Image helpers the alignment path needs (`rgb`, dust3r/utils/image.py:25-40) plus `load_images`
(:74-128), kept so reference scripts run; preprocessing is host-side PIL work outside the hot paths."""
from __future__ import annotations

import os
import numpy as np
import torch


def img_to_arr(img):
    if isinstance(img, str):
        import PIL.Image
        img = np.asarray(PIL.Image.open(img).convert('RGB'))
    return img


def rgb(ftensor, true_shape=None):
    """[-1,1] CHW tensor(s) -> [0,1] HWC numpy."""
    if isinstance(ftensor, list):
        return [rgb(x, true_shape=true_shape) for x in ftensor]
    if isinstance(ftensor, torch.Tensor):
        ftensor = ftensor.detach().cpu().numpy()
    if ftensor.ndim == 3 and ftensor.shape[0] == 3:
        ftensor = ftensor.transpose(1, 2, 0)
    elif ftensor.ndim == 4 and ftensor.shape[1] == 3:
        ftensor = ftensor.transpose(0, 2, 3, 1)
    if true_shape is not None:
        H, W = true_shape
        ftensor = ftensor[:H, :W]
    if ftensor.dtype == np.uint8:
        img = np.float32(ftensor) / 255
    else:
        img = (ftensor * 0.5) + 0.5
    return img.clip(min=0, max=1)


def _resize_pil_image(img, long_edge_size):
    import PIL.Image
    S = max(img.size)
    interp = PIL.Image.LANCZOS if S > long_edge_size else PIL.Image.BICUBIC
    new_size = tuple(int(round(x * long_edge_size / S)) for x in img.size)
    return img.resize(new_size, interp)


def load_images(folder_or_list, size, square_ok=False, verbose=True):
    """Open, resize (long edge -> size; 224 = short edge then centre crop), crop to multiples of 16
    and normalise to [-1,1]; returns the list of view dicts inference() consumes."""
    import PIL.Image
    from PIL.ImageOps import exif_transpose
    if isinstance(folder_or_list, str):
        root, folder_content = folder_or_list, sorted(os.listdir(folder_or_list))
    elif isinstance(folder_or_list, list):
        root, folder_content = '', folder_or_list
    else:
        raise ValueError(f'bad {folder_or_list=} ({type(folder_or_list)})')
    imgs = []
    for path in folder_content:
        if not path.lower().endswith(('.jpg', '.jpeg', '.png')):
            continue
        img = exif_transpose(PIL.Image.open(os.path.join(root, path))).convert('RGB')
        W1, H1 = img.size
        if size == 224:
            img = _resize_pil_image(img, round(size * max(W1 / H1, H1 / W1)))
        else:
            img = _resize_pil_image(img, size)
        W, H = img.size
        cx, cy = W // 2, H // 2
        if size == 224:
            half = min(cx, cy)
            img = img.crop((cx - half, cy - half, cx + half, cy + half))
        else:
            halfw, halfh = ((2 * cx) // 16) * 8, ((2 * cy) // 16) * 8
            if not square_ok and W == H:
                halfh = 3 * halfw / 4
            img = img.crop((cx - halfw, cy - halfh, cx + halfw, cy + halfh))
        arr = torch.from_numpy(np.asarray(img, dtype=np.float32) / 255).permute(2, 0, 1)
        arr = (arr - 0.5) / 0.5
        if verbose:
            print(f' - adding {path} with resolution {W1}x{H1} --> {img.size[0]}x{img.size[1]}')
        imgs.append(dict(img=arr[None], true_shape=np.int32([img.size[::-1]]), idx=len(imgs), instance=str(len(imgs))))
    assert imgs, 'no images found at ' + root
    if verbose:
        print(f' (Found {len(imgs)} images)')
    return imgs
