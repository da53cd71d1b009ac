"""pytest configuration: ``gpu`` marker, repo root on sys.path, shared golden-fixture loader."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN_DIR = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| in float64."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm())


def load_golden(name: str) -> dict:
    """Load a fixture written by oracle/make_golden.py; regenerate seed-derived inputs if not stored."""
    import hashlib
    import b200cam.synthetic as synth
    z = np.load(GOLDEN_DIR / f"{name}.npz")
    out = {k: (torch.from_numpy(z[k]) if z[k].dtype.kind == "f" else z[k]) for k in z.files}
    N, B = int(z["N"][0]), int(z["B"][0])
    out["N"], out["B"] = N, B
    if "img" not in out:
        s_img, s_w, s_h = (int(v) for v in z["seeds"])
        out["img"] = synth.images(B, N, s_img)
        out["w"] = synth.upstream_grad(B, N, s_w)
        out["h"] = synth.height_map(N, s_h)
        m = hashlib.sha256()
        for t in (out["img"], out["w"], out["h"]):
            m.update(t.contiguous().numpy().tobytes())
        if m.hexdigest() != str(z["input_sha256"][0]):
            pytest.skip(f"{name}: torch RNG stream differs from the one the fixture was made with")
    return out


GOLDEN_CASES = ["facedeid_n64_b2", "facedeid_n128_b3", "facedeid_n256_b2"]
