"""CPU restatement (test infrastructure only) of the loaders' point -> image maps
(``lib/dataset/nuscenes_dataloader.py:274-278``: sparse depth and 2D label map) and of the RGB point features
(``:364-367``).  The reference builds them inline inside ``__getitem__`` (which needs the dataset on disk), so the
statements are restated with the same numpy operations rather than imported; numpy's indexed assignment defines the
duplicate-pixel rule (the last point wins) and is what the CUDA kernel is checked against."""
import numpy as np


def rasterize(img_indices, values, height, width, fill):
    """One sample: float64 map like the reference's, [H, W]."""
    out = np.ones((height, width)) * fill                            # nuscenes_dataloader.py:275 / :277
    out[img_indices[:, 0], img_indices[:, 1]] = values               # :276 / :278
    return out


def rgb_feats(img_chw, img_indices):
    """[3, H, W] image -> [N, 3] point colours (nuscenes_dataloader.py:364-367)."""
    return img_chw[:, img_indices[:, 0], img_indices[:, 1]].T


def fliplr(img_indices, maps, width):
    """The loaders' horizontal flip of indices and maps (:296-301)."""
    idx = img_indices.copy()
    idx[:, 1] = width - 1 - idx[:, 1]
    return idx, [np.ascontiguousarray(np.fliplr(m)) for m in maps]
