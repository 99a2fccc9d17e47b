"""CPU: the C-ABI library loads and exports every symbol include/controlar_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "controlar_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(car_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_all_symbols():
    from controlar_b200.build import build
    path = build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"missing export {n}"


def test_ctypes_prototypes_cover_header():
    from controlar_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared()
    l = _lib.lib()
    assert l.car_version() >= 100
    # argument validation happens before any CUDA call
    assert l.car_model_create(None, None, None, None) < 0
    assert b"null" in l.car_last_error()


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from controlar_b200.autoregressive.models.gpt_t2i import GPT_models
    m = GPT_models["GPT-B"](block_size=64, cls_token_num=120, model_type="t2i", vocab_size=2048).eval()
    with pytest.raises(RuntimeError):
        m.setup_caches(2, 184, torch.float32)
