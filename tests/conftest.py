import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def duck_pt():
    import _oracle as O
    import rayfinder_b200 as rf

    return rf.PtFormat.loads(O.duck_pt_bytes())


@pytest.fixture(scope="session")
def sponza_pt():
    from rayfinder_b200 import assets as rfa

    if rfa.scene_path("Sponza") is None:
        message = "assets/Sponza.pt[.xz] not baked (run __graft_entry__.build() where /root/reference is mounted)"
        try:
            import torch

            on_gpu_box = torch.cuda.is_available()
        except Exception:
            on_gpu_box = False
        if on_gpu_box:
            # every BASELINE.json config but the first is a Sponza config: on the GPU box a missing scene is a failure of
            # the run, never a reason to drop those tests silently
            pytest.fail(message)
        pytest.skip(message)
    return rfa.load_scene("Sponza")


@pytest.fixture(scope="session")
def golden():
    import _oracle as O

    return {p.stem: np.load(p) for p in O.GOLDEN.glob("*.npz")}
