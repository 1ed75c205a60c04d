"""CPU oracle for the training-sample assembly (TEST INFRASTRUCTURE -- never imported by the product).

numpy restatement of basicsr/data/image_npy_dataset.py:189-232 for the `one_voxel_flg` / `return_deblur_voxel`
configuration every blurry option file uses: `triple_random_crop` (transforms.py:211-231) with given top / left,
`augment`'s `_augment` (transforms.py:114-129) with given flags, `img2tensor` (img_util.py:22-28: BGR->RGB for 3-channel
images, HWC->CHW), the deblur-voxel channel packing (:209-221) and the sliding two-bin windows (:226-232).
Pinned by tests/golden/sample_pack_cases.npz (tests/golden/make_sample_golden.py calls the unmodified reference functions).
"""
import numpy as np


def assemble(img_lqs_bgr, img_gts_bgr, voxel_hwc, m, n, gt_size, top, left, hflip, vflip, rot90):
    """img_*_bgr: lists of (H,W,3) float32 BGR arrays as `imfrombytes` returns them; voxel_hwc (H,W,num_bins)."""
    def crop(a):
        return a if gt_size is None else a[top:top + gt_size, left:left + gt_size, ...]

    def aug(a):
        a = np.float32(a)
        if hflip:
            a = a[:, ::-1]
        if vflip:
            a = a[::-1]
        if rot90:
            a = a.transpose(1, 0, 2)
        return a

    def totensor(a):
        if a.shape[2] == 3:
            a = a[:, :, ::-1]  # BGR -> RGB
        return np.ascontiguousarray(a.transpose(2, 0, 1), dtype=np.float32)

    lqs = np.stack([totensor(aug(crop(a))) for a in img_lqs_bgr])
    gts = np.stack([totensor(aug(crop(a))) for a in img_gts_bgr])
    vox = totensor(aug(crop(voxel_hwc)))                          # (num_bins,h,w)
    lq = np.concatenate((lqs[0], vox[1:m], lqs[1], vox[m + 2 + n:]), 0)
    win = np.stack([vox[i:i + 2] for i in range(vox.shape[0] - 1)])
    return {"lq": lq, "voxel": win, "gt": gts}
