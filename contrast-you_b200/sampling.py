"""Dense-feature point sampling for the dense InfoNCE hook (reference: semi_seg/hooks/infonce.py:31-46).

Same coordinates as the reference for a given seed: per image, ``point_nums`` distinct rows and columns drawn from the
legacy numpy generator seeded with ``seed`` (the reference seeds the *global* numpy state inside a context manager
that restores it afterwards, contrastyou/utils/utils.py:121-141; a private RandomState(seed) yields the identical
stream without touching global state).  The gather itself is one advanced-indexing call on the device instead of
B x point_nums slices and a ``torch.stack``.
"""
import numpy as np
import torch

__all__ = ["region_coordinates", "region_extractor", "region_pixel_offsets"]


def region_coordinates(batch: int, h: int, w: int, point_nums: int, seed: int) -> np.ndarray:
    rs = np.random.RandomState(seed)
    out = np.empty((batch, point_nums, 2), dtype=np.int64)
    for b in range(batch):
        out[b, :, 0] = rs.choice(range(h), point_nums, replace=False)
        out[b, :, 1] = rs.choice(range(w), point_nums, replace=False)
    return out


def region_extractor(normalize_features: torch.Tensor, *, point_nums=5, seed: int) -> torch.Tensor:
    """[B, C, h, w] -> [B * point_nums, C], image-major, in the reference's point order."""
    b, c, h, w = normalize_features.shape
    coords = torch.from_numpy(region_coordinates(b, h, w, point_nums, seed)).to(normalize_features.device)
    bi = torch.arange(b, device=normalize_features.device)[:, None].expand(b, point_nums)
    picked = normalize_features[bi, :, coords[..., 0], coords[..., 1]]       # [B, P, C]
    return picked.reshape(b * point_nums, c)


def region_pixel_offsets(batch: int, channels: int, h: int, w: int, point_nums: int, seed: int) -> torch.Tensor:
    """element offsets b*C*h*w + y*w + x of the hook's sampled pixels inside a contiguous [B, C, h, w] map, image-major in the
    reference's point order — the pixel list ``SupConLoss1.forward_dense`` gathers inside its pack kernel"""
    coords = region_coordinates(batch, h, w, point_nums, seed)                  # [B, P, (y, x)]
    base = np.arange(batch, dtype=np.int64)[:, None] * (channels * h * w)
    return torch.from_numpy((base + coords[..., 0] * w + coords[..., 1]).reshape(-1))
