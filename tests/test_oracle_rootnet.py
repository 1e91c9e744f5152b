"""oracle.rootnet (RootNet / SAR depth) against the fixture produced by the reference's own classes and
preprocessing functions (tests/golden/make_golden_rootnet.py)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import rootnet as orn

import importlib.util
_spec = importlib.util.spec_from_file_location('make_golden_rootnet', os.path.join(GOLDEN, 'make_golden_rootnet.py'))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


@pytest.fixture(scope='module')
def model():
    return orn.seeded_rootnet(0)


def test_rootnet_oracle_reproduces_the_reference(model):
    z = np.load(os.path.join(GOLDEN, 'rootnet_seed0.npz'))
    for i, (h, w, seed, box, fx, fy) in enumerate(mg.CASES):
        img = mg.synthetic_frame(h, w, seed)
        patch, bbox = orn.crop_patch(img, box)
        assert np.array_equal(bbox, z[f'bbox{i}'])
        assert hashlib.sha256(patch.numpy().tobytes()).digest() == z[f'patch_sha{i}'].tobytes()      # crop bit-exact
        K = np.array([[fx, 0, w / 2], [0, fy, h / 2], [0, 0, 1]])
        depth = model.estimate_root_depth_custom(img, K, box)
        assert depth == pytest.approx(float(z[f'depth{i}']), rel=1e-6, abs=1e-7)


def test_degenerate_box_is_rejected():
    assert orn.process_bbox([10, 10, 0, 50], 640, 480) is None
