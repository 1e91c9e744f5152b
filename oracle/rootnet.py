"""Oracle: RootNet / SAR root-depth estimate (numpy + torch on CPU).  TEST INFRASTRUCTURE ONLY.

Restates EstimateRGB.estimate_root_depth_custom (rootnet/Model_RGB.py:572-638), the call infer.py makes per hand right
after the HaMeR forward (hamer/infer.py:1275):
  rootnet/preprocessing.py:154-165   sanitize_bbox
  rootnet/preprocessing.py:167-188   process_bbox(bbox_xywh, W, H, (256, 256), ratio 1.5)
  rootnet/preprocessing.py:39-79     generate_patch_image (gen_trans_from_patch_cv :102-152, cv2.warpAffine with the float32 matrix)
  rootnet/Model_RGB.py:317-318       transform = ToTensor + Normalize(ImageNet mean / std)   (torchvision)
  rootnet/Model_RGB.py:179-197       SARresnet34 = torchvision resnet34 stem + layer1..4 (third-party torchvision.models.resnet34,
                                     torchvision 0.26 in this image: instantiated here exactly like the reference does)
  rootnet/Model_RGB.py:240-287       ResRootNet: global average pool -> 1x1 conv (512 -> 1) -> gamma * k_value
  rootnet/Model_RGB.py:494-498       calculate_k = sqrt(0.3 * 0.3 * fx * fy / (bbox_w * bbox_h))
Pinned: tests/golden/rootnet_seed0.npz holds outputs of the reference's own SARresnet34 / ResRootNet classes and
preprocessing functions (tests/golden/make_golden_rootnet.py executes them from /root/reference).
"""
import numpy as np
import torch
import torch.nn as nn

from . import crop as oc

INPUT_SHAPE = (256, 256)          # sar_config_stage_1.py:17
BBOX_REAL = (0.3, 0.3)            # sar_config_stage_1.py:19
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def sanitize_bbox(bbox, img_width, img_height):
    """preprocessing.py:154-165."""
    x, y, w, h = bbox
    x1 = np.max((0, x))
    y1 = np.max((0, y))
    x2 = np.min((img_width - 1, x1 + np.max((0, w - 1))))
    y2 = np.min((img_height - 1, y1 + np.max((0, h - 1))))
    if w * h > 0 and x2 > x1 and y2 > y1:
        return np.array([x1, y1, x2 - x1, y2 - y1])
    return None


def process_bbox(bbox, img_width, img_height, input_img_shape=INPUT_SHAPE, ratio=1.5):
    """preprocessing.py:167-188 (float64 arithmetic, float32 result)."""
    bbox = sanitize_bbox(bbox, img_width, img_height)
    if bbox is None:
        return bbox
    w, h = bbox[2], bbox[3]
    c_x, c_y = bbox[0] + w / 2., bbox[1] + h / 2.
    aspect_ratio = input_img_shape[1] / input_img_shape[0]
    if w > aspect_ratio * h:
        h = w / aspect_ratio
    elif w < aspect_ratio * h:
        w = h * aspect_ratio
    bbox[2] = w * ratio
    bbox[3] = h * ratio
    bbox[0] = c_x - bbox[2] / 2.
    bbox[1] = c_y - bbox[3] / 2.
    return bbox.astype(np.float32)


def patch_affine(bbox):
    """generate_patch_image's matrix (preprocessing.py:62-72, :102-152 with scale 1, rot 0): float32 source points, LU solve in
    double (cv2.getAffineTransform), result cast to float32 - and that float32 matrix is what cv2.warpAffine receives."""
    bb_c_x = float(bbox[0] + 0.5 * bbox[2])
    bb_c_y = float(bbox[1] + 0.5 * bbox[3])
    bb_w, bb_h = float(bbox[2]), float(bbox[3])
    src = np.zeros((3, 2), np.float32)
    src[0] = (bb_c_x, bb_c_y)
    src[1] = src[0] + np.array([0, bb_h * 0.5], np.float32)
    src[2] = src[0] + np.array([bb_w * 0.5, 0], np.float32)
    dst = np.array([[128, 128], [128, 256], [256, 128]], np.float32)
    A, b = [], []
    for (sx, sy), (dx, dy) in zip(src.astype(np.float64), dst.astype(np.float64)):
        A.append([sx, sy, 1, 0, 0, 0]); b.append(dx)
        A.append([0, 0, 0, sx, sy, 1]); b.append(dy)
    M = np.array(oc._lu_solve6(A, b), dtype=np.float64).reshape(2, 3)
    return M.astype(np.float32)


def crop_patch(img_bgr, bbox_xyxy):
    """Model_RGB.py:587-612: box -> processed bbox -> 256 x 256 patch -> RGB, ToTensor, Normalize.  Returns
    (tensor [3,256,256] float32, processed bbox float32 [4])."""
    x1, y1, x2, y2 = bbox_xyxy
    height, width = img_bgr.shape[:2]
    bbox = process_bbox([x1, y1, x2 - x1, y2 - y1], width, height)
    M = patch_affine(bbox)
    patch = oc.warp_affine_u8(img_bgr, M.astype(np.float64))            # cv2.warpAffine(img, float32 M, (256, 256), INTER_LINEAR)
    patch = patch.astype(np.float32)[:, :, ::-1].astype(np.uint8)
    t = torch.from_numpy(np.ascontiguousarray(patch.transpose(2, 0, 1))).to(torch.float32).div(255)   # ToTensor
    mean = torch.as_tensor(MEAN, dtype=torch.float32)[:, None, None]
    std = torch.as_tensor(STD, dtype=torch.float32)[:, None, None]
    return (t - mean) / std, bbox


class SARresnet34(nn.Module):
    """Model_RGB.py:179-197."""

    def __init__(self):
        super().__init__()
        from torchvision import models
        backbone = models.resnet34(weights=None)
        self.extract_mid = nn.Sequential(backbone.conv1, backbone.bn1, backbone.relu, backbone.maxpool, backbone.layer1, backbone.layer2)
        self.extract_high = nn.ModuleList([nn.Sequential(backbone.layer3, backbone.layer4)])

    def forward(self, x):
        return self.extract_high[0](self.extract_mid(x))


class ResRootNet(nn.Module):
    """Model_RGB.py:240-287 (inference branch)."""

    def __init__(self, inplanes=512):
        super().__init__()
        self.depth_layer = nn.Conv2d(inplanes, 1, kernel_size=1, stride=1, padding=0)

    def forward(self, x, k_value):
        img_feat = torch.mean(x.view(x.size(0), x.size(1), x.size(2) * x.size(3)), dim=2)
        gamma = self.depth_layer(img_feat[:, :, None, None]).view(-1, 1)
        return gamma * k_value.view(-1, 1)


def calculate_k(bbox, fx, fy):
    """Model_RGB.py:494-498 (float32 tensor arithmetic: tensor(0.09) * fx * fy / area, sqrt)."""
    area = bbox[-1] * bbox[-2]
    real_area = torch.tensor(BBOX_REAL[0] * BBOX_REAL[1])
    return torch.sqrt(real_area * fx * fy / area).unsqueeze(0)


class ORootNet(nn.Module):
    """backbone + rootnet with the reference's state-dict layout: 'backbone.*' (checkpoint['model'] restricted to the
    backbone) and 'rootnet.depth_layer.*' (checkpoint['rootnet'])."""

    def __init__(self):
        super().__init__()
        self.backbone = SARresnet34()
        self.rootnet = ResRootNet(512)

    @torch.no_grad()
    def estimate_root_depth_custom(self, img_bgr, K, bbox_xyxy):
        patch, bbox = crop_patch(img_bgr, bbox_xyxy)
        feats = self.backbone(patch[None])
        K = np.asarray(K)
        k_value = calculate_k(bbox, K[0, 0], K[1, 1])
        return self.rootnet(feats, k_value).item()


def seeded_rootnet(seed=0):
    """Random-init model the way the constructors leave it (torchvision's kaiming fan_out init), with BatchNorm statistics and
    affine parameters perturbed so that the BN folding is exercised; >= 2-D weights rounded to bf16 like the HaMeR fixture."""
    torch.manual_seed(seed)
    m = ORootNet().eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
                mod.weight.copy_(torch.rand(mod.num_features, generator=g) * 0.4 + 0.5)
                mod.bias.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
    return m
