"""Golden fixture for the RootNet / SAR depth path: outputs of the REFERENCE's own code, generated in the
authoring container (needs /root/reference; the fixture travels, the reference does not).

    python tests/golden/make_golden_rootnet.py

rootnet/Model_RGB.py cannot be imported as a module (it pulls convnext, MANO pickles, onnx ...), so the two classes on the
path - SARresnet34 (Model_RGB.py:179-197) and ResRootNet (:240-302) - are cut out of its source by ast and executed
unmodified; rootnet/preprocessing.py imports cleanly and provides process_bbox / generate_patch_image.  Stored:
per-case processed bbox, a hash + samples of the normalised patch, the backbone feature checksum and the predicted depth,
for the seeded model of oracle.rootnet.seeded_rootnet (same parameter tree: asserted here).
"""
import ast
import hashlib
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference/rootnet'


def reference_classes():
    src = open(os.path.join(REF, 'Model_RGB.py')).read()
    tree = ast.parse(src)
    want = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ('SARresnet34', 'ResRootNet')]
    assert len(want) == 2
    from torchvision import models
    ns = {'nn': nn, 'torch': torch, 'F': F, 'models': models}
    exec(compile(ast.Module(body=want, type_ignores=[]), 'Model_RGB.py', 'exec'), ns)
    return ns['SARresnet34'], ns['ResRootNet']


def reference_preprocessing():
    import types
    ply = types.ModuleType('plyfile')            # imported at module level for a point-cloud writer that is not on this path
    ply.PlyData = ply.PlyElement = None
    sys.modules.setdefault('plyfile', ply)
    spec = importlib.util.spec_from_file_location('ref_preprocessing', os.path.join(REF, 'preprocessing.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synthetic_frame(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) * 255 // (w + h))], axis=-1)
    return np.clip(base + rng.integers(-40, 40, (h, w, 3)), 0, 255).astype(np.uint8)


CASES = [  # (frame h, w, seed, bbox xyxy, fx, fy)
    (480, 640, 41, [100, 80, 260, 300], 600.0, 610.0),
    (480, 640, 41, [330.5, 120.25, 520, 420], 600.0, 610.0),
    (1080, 1920, 42, [1500, 700, 1919, 1079], 906.96, 906.79),
    (720, 1280, 43, [-20, 10, 180, 260], 900.0, 900.0),
]


def main():
    import torchvision.transforms as standard
    from oracle import rootnet as orn
    SAR, RRN = reference_classes()
    pre = reference_preprocessing()
    ora = orn.seeded_rootnet(0)
    ref_b, ref_r = SAR().eval(), RRN(inplanes=512).eval()
    assert [k for k, _ in ref_b.state_dict().items()] == [k for k, _ in ora.backbone.state_dict().items()]
    ref_b.load_state_dict(ora.backbone.state_dict())
    ref_r.load_state_dict(ora.rootnet.state_dict())
    transform = standard.Compose([standard.ToTensor(), standard.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
    out = {}
    for i, (h, w, seed, box, fx, fy) in enumerate(CASES):
        img = synthetic_frame(h, w, seed)
        x1, y1, x2, y2 = box
        bbox = pre.process_bbox([x1, y1, x2 - x1, y2 - y1], w, h, (256, 256), 1.5)           # Model_RGB.py:587-595
        patch, _, _ = pre.generate_patch_image(cvimg=img, bbox=bbox, scale=1.0, rot=0.0, do_flip=False, out_shape=(256, 256))
        patch = patch[:, :, ::-1].astype(np.uint8).copy()
        t = transform(patch)
        with torch.no_grad():
            feats = ref_b(t[None])
            area = bbox[-1] * bbox[-2]
            k_value = torch.sqrt(torch.tensor(0.3 * 0.3) * fx * fy / area).unsqueeze(0)           # calculate_k, Model_RGB.py:494-498
            depth = ref_r(feats, k_value).item()
        out[f'bbox{i}'] = bbox
        out[f'patch_sha{i}'] = np.frombuffer(hashlib.sha256(t.numpy().tobytes()).digest(), np.uint8)
        out[f'patch_samples{i}'] = t.numpy()[:, ::37, ::41].copy()
        out[f'feat_mean{i}'] = feats.mean(dim=(2, 3))[0].numpy()
        out[f'depth{i}'] = np.float64(depth)
        print(i, bbox, depth, float(feats.abs().max()))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'rootnet_seed0.npz'), **out)


if __name__ == '__main__':
    main()
