"""Image I/O on the edges of the two hot paths: `load_images` builds the view dicts `inference()` consumes and `rgb`
turns normalised tensors back into displayable arrays for the optimizer's `imgs` attribute.  Host-side PIL work; the
conventions (long edge -> `size`, dimensions cropped to multiples of 16, 4:3 crop of square inputs, [-1, 1] range,
`true_shape`, `idx`, `instance`) are those of dust3r/utils/image.py:74-128."""
from __future__ import annotations

import os

import numpy as np
import torch

_EXTENSIONS = ('.jpg', '.jpeg', '.png')


def rgb(ftensor, true_shape=None):
    """Normalised image(s) -> float arrays in [0, 1], channels last.  Accepts a tensor / array (CHW, BCHW or already
    channels-last), uint8 or [-1, 1] float data, or a list of those; `true_shape` = (H, W) crops the result."""
    if isinstance(ftensor, list):
        return [rgb(item, true_shape=true_shape) for item in ftensor]
    arr = ftensor.detach().cpu().numpy() if torch.is_tensor(ftensor) else np.asarray(ftensor)
    if arr.ndim == 3 and arr.shape[0] == 3:
        arr = np.moveaxis(arr, 0, -1)
    elif arr.ndim == 4 and arr.shape[1] == 3:
        arr = np.moveaxis(arr, 1, -1)
    if true_shape is not None:
        height, width = true_shape
        arr = arr[:height, :width]
    arr = arr.astype(np.float32) / 255 if arr.dtype == np.uint8 else arr * 0.5 + 0.5
    return arr.clip(min=0, max=1)


def img_to_arr(img):
    """A path is opened as RGB uint8; arrays pass through."""
    if isinstance(img, str):
        import PIL.Image
        return np.asarray(PIL.Image.open(img).convert('RGB'))
    return img


def _rescale(img, long_edge):
    """Resize so that the long edge becomes `long_edge` (Lanczos when shrinking, bicubic when enlarging)."""
    import PIL.Image
    current = max(img.size)
    method = PIL.Image.LANCZOS if current > long_edge else PIL.Image.BICUBIC
    return img.resize(tuple(int(round(side * long_edge / current)) for side in img.size), method)


def _crop_box(width, height, size, square_ok):
    """Centre crop: a square for the 224 models, otherwise both sides rounded down to multiples of 16 (and a square
    image cut to 4:3 unless square_ok)."""
    cx, cy = width // 2, height // 2
    if size == 224:
        half_w = half_h = min(cx, cy)
    else:
        half_w, half_h = ((2 * cx) // 16) * 8, ((2 * cy) // 16) * 8
        if width == height and not square_ok:
            half_h = 3 * half_w / 4
    return (cx - half_w, cy - half_h, cx + half_w, cy + half_h)


def load_images(folder_or_list, size, square_ok=False, verbose=True):
    """Folder name or list of file names -> list of dict(img (1,3,H,W) in [-1,1], true_shape int32 [[H,W]], idx,
    instance) ready for make_pairs / inference.  Files that are not .jpg/.jpeg/.png are skipped."""
    import PIL.Image
    from PIL.ImageOps import exif_transpose
    if isinstance(folder_or_list, str):
        root, names = folder_or_list, sorted(os.listdir(folder_or_list))
    elif isinstance(folder_or_list, list):
        root, names = '', folder_or_list
    else:
        raise ValueError(f'bad {folder_or_list=} ({type(folder_or_list)})')
    views = []
    for name in names:
        if not name.lower().endswith(_EXTENSIONS):
            continue
        pil = exif_transpose(PIL.Image.open(os.path.join(root, name))).convert('RGB')
        w_in, h_in = pil.size
        # 224 models: the SHORT edge becomes 224 (then a square crop); the others: the long edge becomes `size`
        long_edge = round(size * max(w_in / h_in, h_in / w_in)) if size == 224 else size
        pil = _rescale(pil, long_edge)
        pil = pil.crop(_crop_box(pil.size[0], pil.size[1], size, square_ok))
        pixels = torch.from_numpy(np.asarray(pil, dtype=np.float32) / 255).permute(2, 0, 1)
        if verbose:
            print(f' - adding {name} with resolution {w_in}x{h_in} --> {pil.size[0]}x{pil.size[1]}')
        views.append(dict(img=((pixels - 0.5) / 0.5)[None], true_shape=np.int32([pil.size[::-1]]), idx=len(views),
                          instance=str(len(views))))
    assert views, 'no images found at ' + root
    if verbose:
        print(f' (Found {len(views)} images)')
    return views
