"""Token-selection visualiser: host-side mirror of the reference's
`projects/mmdet3d_plugin/models/utils/token_select_vis.py:8-79`, the only in-tree consumer of the backbone's
`token_masks` / `keep_idx` / `drop_idx` (called from `detectors/petr3d.py:562-579` when `token_select_vis=True`).

CPU / numpy code by nature (it writes PNG files); no kernel is involved.  `token_selection_overlays` returns the very
arrays the reference hands to `mmcv.imwrite` (name -> H x W x 3 or H x W x 4 float32), `token_selection_vis` writes them.
mmcv (absent from this image, pinned by the reference to mmcv-full 1.6.0) is restated through the OpenCV calls it
makes: `imdenormalize` = cv2.multiply(img, std64) ; cv2.add(img, mean64) ; `imwrite` = cv2.imwrite after makedirs.
"""
import os

import numpy as np


def _imdenormalize(img, mean, std):
    """mmcv.imdenormalize(img, mean, std, to_bgr=False) (mmcv/image/photometric.py)."""
    import cv2
    assert img.dtype != np.uint8
    mean = np.asarray(mean, dtype=np.float32).reshape(1, -1).astype(np.float64)
    std = np.asarray(std, dtype=np.float32).reshape(1, -1).astype(np.float64)
    img = cv2.multiply(np.ascontiguousarray(img), std)      # makes a copy
    cv2.add(img, mean, img)
    return img


def _to_numpy(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def token_selection_overlays(input_imgs, masks, keep_idxes=None, drop_idxes=None, img_norm_cfg=None, patch_size=16,
                             min_alpha=0.3, max_alpha=1.0):
    """token_select_vis.py:27-79 without the file writes.  input_imgs (views, 3, H, W) normalised images; masks: list
    of (views, H/ps, W/ps, 1); keep_idxes / drop_idxes: lists of (views, k) / (views, N - k) int64 (both or neither).
    -> dict file name (as the reference names it) -> float32 array."""
    imgs = _to_numpy(input_imgs)
    masks = [_to_numpy(m) for m in masks]
    both = keep_idxes is not None and drop_idxes is not None
    if both:
        keep_idxes = [_to_numpy(k) for k in keep_idxes]
    out = {}
    for v in range(imgs.shape[0]):
        img = np.transpose(imgs[v], [1, 2, 0])
        if img_norm_cfg is not None:
            img = _imdenormalize(img, img_norm_cfg["mean"], img_norm_cfg["std"])
        alpha = np.ones([*img.shape[:2], 1], dtype=img.dtype) * 255
        for l, layer_mask in enumerate(masks):
            m = layer_mask[v]
            if img.shape[0] // m.shape[0] != patch_size or img.shape[1] // m.shape[1] != patch_size:
                raise AssertionError("mask grid does not match the image / patch size (token_select_vis.py:47-48)")
            px = np.repeat(np.repeat(m, patch_size, axis=1), patch_size, axis=0)
            px = px * (max_alpha - min_alpha) + min_alpha
            out["view%d_layer%d.png" % (v, l)] = np.concatenate([img, alpha * px], axis=-1)
            out["view%d_layer%d_ori.png" % (v, l)] = img
        if both:
            nh, nw = img.shape[0] // patch_size, img.shape[1] // patch_size
            for l, layer_keep in enumerate(keep_idxes):
                k = layer_keep[v]
                if k.max() >= nh * nw:
                    raise AssertionError("keep index beyond the patch grid (token_select_vis.py:67)")
                km = np.zeros([nh * nw], dtype=img.dtype)
                km[k] = 1
                km = km.reshape(nh, nw, 1)
                km = np.repeat(np.repeat(km, patch_size, axis=1), patch_size, axis=0)
                km = km * (max_alpha - min_alpha) + min_alpha
                out["view%d_layer%d_keepidx.png" % (v, l)] = np.concatenate([img, alpha * km], axis=-1)
    return out


def token_selection_vis(input_imgs, masks, keep_idxes, drop_idxes, img_norm_cfg, output_path, patch_size=16,
                        min_alpha=0.3, max_alpha=1.0):
    """Same signature as the reference function; writes the PNGs under output_path."""
    import cv2
    os.makedirs(output_path, exist_ok=True)
    for name, arr in token_selection_overlays(input_imgs, masks, keep_idxes, drop_idxes, img_norm_cfg, patch_size,
                                              min_alpha, max_alpha).items():
        cv2.imwrite(os.path.join(output_path, name), arr)
