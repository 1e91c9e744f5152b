import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = '/root/reference'          # only present in the authoring container
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box')


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped (not failed) when no CUDA device is visible, e.g. `pytest tests/` here
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def oracle_model():
    """The fp32 CPU oracle with the seed-0 random init of the reference constructors and
    bf16-representable >=2-D weights.  ~15 s to build."""
    import torch
    from oracle.pipeline import OHamer
    from hamer_yolo_b200.mano_data import synthetic_mano
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = OHamer(synthetic_mano(0))
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.copy_(p.to(torch.bfloat16).float())
    return model


@pytest.fixture(scope='session')
def seeded_images():
    import torch
    g = torch.Generator().manual_seed(1234)
    return torch.randn(2, 3, 256, 256, generator=g).to(torch.bfloat16).float()
