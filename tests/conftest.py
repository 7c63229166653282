import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))
    return {k: torch.from_numpy(z[k].astype(np.int64) if z[k].dtype == np.int16 else z[k]) for k in z.files}


@pytest.fixture(scope="session")
def gin():
    """The seeded inputs the golden outputs were produced from."""
    from tests.golden.make_golden import inputs_ops
    return inputs_ops()


@pytest.fixture(scope="session")
def golden_refnerf():
    """Outputs of the unmodified reference's Ref-NeRF helpers (tests/golden/make_golden.py refnerf)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs_refnerf.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def golden_ref_train():
    """Training side of Ref-NeRF from the unmodified reference (tests/golden/make_golden.py round3): RefNeRF.get_grad,
    the normal / back-face losses and every parameter gradient of one is_ref_model loss (train.py:176-199)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs_ref_train.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def golden_round2():
    """Round-2 outputs of the unmodified reference (tests/golden/make_golden.py round2): seeded multi-tile render_image,
    Ref-NeRF forward / Ref branch of render_image, one training step's losses and gradient summaries."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs_round2.npz"))
    return {k: torch.from_numpy(z[k].astype(np.int64) if z[k].dtype == np.int16 else z[k]) for k in z.files}
