"""CPU: the Canny restatement (oracle/canny_oracle.py) against the fixture that cv2.Canny itself produced (tests/golden/canny.npz,
tests/golden/make_golden.py:canny_case — the call of reference condition/canny.py:14) and, where OpenCV is importable, against
cv2.Canny live; the left-padding restatement against the reference's own lines."""
import numpy as np
import pytest
import torch

from oracle.canny_oracle import canny, left_pad_captions
from tests.golden.make_golden import canny_inputs
from tests.helpers import GOLDEN

PAIRS = ((100, 200), (50, 150), (30.5, 90.7))


def test_canny_oracle_matches_cv2_fixture_bit_exactly():
    g = np.load(f"{GOLDEN}/canny.npz")
    n = 0
    for name, img in canny_inputs().items():
        for lo, hi in PAIRS:
            want = g[f"{name}_{lo}_{hi}"]
            got = canny(img, lo, hi)
            assert got.dtype == np.uint8 and got.shape == want.shape
            assert np.array_equal(got, want), (name, lo, hi, int((got != want).sum()))
            n += 1
    assert n == 12


def test_canny_oracle_matches_cv2_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for t in range(6):
        H, W = int(rng.integers(16, 120)), int(rng.integers(16, 120))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        if t % 2:
            img = cv2.GaussianBlur(img, (7, 7), 0)
        assert np.array_equal(canny(img, 100, 200), cv2.Canny(img, 100, 200))
        assert np.array_equal(canny(img, 200, 100), cv2.Canny(img, 200, 100))       # swapped thresholds


def test_left_pad_oracle_semantics():
    emb = torch.arange(2 * 5 * 3, dtype=torch.float32).reshape(2, 5, 3)
    mask = torch.tensor([[1, 1, 0, 0, 0], [1, 1, 1, 1, 0]])
    new, nm = left_pad_captions(emb, mask)
    assert nm.tolist() == [[0, 0, 0, 1, 1], [0, 1, 1, 1, 1]]
    assert torch.equal(new[0, 3:], emb[0, :2]) and torch.equal(new[0, :3], emb[0, 2:])
    assert torch.equal(new[1, 1:], emb[1, :4]) and torch.equal(new[1, :1], emb[1, 4:])


def test_hed_module_surface_matches_reference_keys():
    """`controlar_b200.condition.hed.ControlNetHED_Apache2` loads a state dict with the reference's keys (condition/hed.py:37-45)."""
    from controlar_b200.condition.hed import ControlNetHED_Apache2, HEDdetector
    from oracle.weights import make_hed_state_dict
    sd = make_hed_state_dict(seed=1)
    m = ControlNetHED_Apache2()
    assert set(m.state_dict().keys()) == set(sd.keys()) and len(sd) == 1 + 2 * 13 + 2 * 5
    m.load_state_dict(sd, strict=True)
    assert len(m._tensors()) == 37
    with pytest.raises(RuntimeError):
        m.run(torch.zeros(1, 3, 32, 32))                      # CPU tensors: no fallback path
    assert isinstance(HEDdetector().netNetwork, ControlNetHED_Apache2)
