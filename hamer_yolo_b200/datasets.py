"""Drop-in mirror of the upstream-HaMeR dataset wrapper ``ViTDetDataset``
(hamer/hamer/datasets/vitdet_dataset.py:16-95), backed by the same crop / warp / normalise kernel as
``hamer_inference.prepare_batch_bbox``.

Differences from ``prepare_batch_bbox`` that this class reproduces (they change pixels):
  * centre, scale and box size are computed in float32 numpy arithmetic (vitdet_dataset.py:37-42, 52-58);
  * left hands are flipped INSIDE the warp: ``generate_image_patch_cv2(..., do_flip=True)`` mirrors the frame and uses
    ``c_x = img_width - c_x - 1`` (hamer/hamer/datasets/utils.py:347-349) - one output column away from flipping the
    patch afterwards;
  * boxes whose crop window exceeds 2.2 x the patch (bbox_size > 563 px) are anti-aliased with a gaussian blur of the
    WHOLE frame first (vitdet_dataset.py:62-71, skimage.filters.gaussian) and the float64 result is warped: hm_preprocess_blur
    evaluates the same blur at the taps the warp reads, with scipy's arithmetic (csrc/crop_warp.cuh).

(In the reference fork the class itself is broken - it unpacks two values from the three-value return of
generate_image_patch_cv2, vitdet_dataset.py:74 vs datasets/utils.py:376 - so this follows what the code computes up to
that line and what upstream HaMeR returns.)
"""
import numpy as np
import torch

DEFAULT_MEAN = 255. * np.array([0.485, 0.456, 0.406])
DEFAULT_STD = 255. * np.array([0.229, 0.224, 0.225])


def expand_to_aspect_ratio(input_shape, target_aspect_ratio=None):
    """hamer/hamer/datasets/utils.py:15-34, on whatever dtype comes in (ViTDetDataset passes float32)."""
    if target_aspect_ratio is None:
        return input_shape
    try:
        w, h = input_shape
    except (ValueError, TypeError):
        return input_shape
    w_t, h_t = target_aspect_ratio
    if h / w < h_t / w_t:
        h_new = max(w * h_t / w_t, h)
        w_new = w
    else:
        h_new = h
        w_new = max(h * w_t / h_t, w)
    return np.array([w_new, h_new])


class ViTDetDataset(torch.utils.data.Dataset):
    """``ViTDetDataset(cfg, img_cv2, boxes, right, rescale_factor=2.5)``; items are the reference's dicts (numpy).
    ``engine`` is the hamer_yolo_b200 Engine that runs the crop kernel (``model.engine``).  ``batch()`` returns the whole
    set as one device-resident batch (what a DataLoader + ``recursive_to`` would produce) without the per-item copies."""

    def __init__(self, cfg, img_cv2, boxes, right, rescale_factor=2.5, train=False, engine=None, **kwargs):
        super().__init__()
        assert train is False, 'ViTDetDataset is only for inference'
        if engine is None:
            raise ValueError('ViTDetDataset needs engine=<hamer_yolo_b200 Engine> (model.engine): the crops are made on the GPU')
        self.cfg = cfg
        self.img_cv2 = img_cv2
        self.engine = engine
        self.train = train
        self.img_size = cfg.MODEL.IMAGE_SIZE
        self.mean = 255. * np.array(self.cfg.MODEL.IMAGE_MEAN)
        self.std = 255. * np.array(self.cfg.MODEL.IMAGE_STD)
        boxes = np.asarray(boxes).astype(np.float32)
        self.center = (boxes[:, 2:4] + boxes[:, 0:2]) / 2.0
        self.scale = rescale_factor * (boxes[:, 2:4] - boxes[:, 0:2]) / 200.0
        self.personid = np.arange(len(boxes), dtype=np.int32)
        self.right = np.asarray(right).astype(np.float32)

    def __len__(self):
        return len(self.personid)

    def _window(self, idx):
        """(c_x, c_y, bbox_size, flip) of item idx with the reference's arithmetic (vitdet_dataset.py:50-60)."""
        center = self.center[idx].copy()
        scale = self.scale[idx]
        bbox_shape = self.cfg.MODEL.get('BBOX_SHAPE', None)
        bbox_size = expand_to_aspect_ratio(scale * 200, target_aspect_ratio=bbox_shape).max()
        flip = bool(self.right[idx] == 0)
        return center[0], center[1], bbox_size, flip

    def _crops(self, indices):
        frame = self.img_cv2
        W = frame.shape[1]
        ccs = np.empty((len(indices), 3), np.float64)
        modes = np.zeros(len(indices), np.uint8)
        sigmas = [0.0] * len(indices)          # kept as computed (numpy float32): scipy's kernel depends on the type
        sizes = []
        for n, idx in enumerate(indices):
            c_x, c_y, bbox_size, flip = self._window(idx)
            downsampling_factor = (bbox_size * 1.0) / self.img_size / 2.0
            if downsampling_factor > 1.1:                        # vitdet_dataset.py:62-71: gaussian(cvimg, sigma=(f - 1) / 2)
                sigmas[n] = (downsampling_factor - 1) / 2
            if flip:
                c_x = W - c_x - 1                                 # datasets/utils.py:347-349
                modes[n] = 2
            ccs[n] = (float(c_x), float(c_y), float(bbox_size))
            sizes.append(bbox_size)
        img, trans = self.engine.preprocess_centers(frame, ccs, modes, blur_sigma=sigmas)
        return img, sizes

    def __getitem__(self, idx):
        img, sizes = self._crops([idx])
        return {
            'img': img[0].cpu().numpy(),
            'personid': int(self.personid[idx]),
            'box_center': self.center[idx].copy(),
            'box_size': sizes[0],
            'img_size': 1.0 * np.array([self.img_cv2.shape[1], self.img_cv2.shape[0]]),
            'right': self.right[idx].copy(),
        }

    def batch(self):
        """All items as one collated batch with 'img' on the engine's device (torch tensors, DataLoader collation dtypes)."""
        n = len(self)
        img, sizes = self._crops(list(range(n)))
        H, W = self.img_cv2.shape[:2]
        return {
            'img': img,
            'personid': torch.from_numpy(self.personid.astype(np.int64)),
            'box_center': torch.from_numpy(self.center.copy()),
            'box_size': torch.tensor(np.asarray(sizes, dtype=np.float64)),
            'img_size': torch.from_numpy(np.tile(1.0 * np.array([W, H]), (n, 1))),
            'right': torch.from_numpy(self.right.copy()),
        }
