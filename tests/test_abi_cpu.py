"""CPU: the C-ABI library loads (no GPU needed) and exports every symbol include/spe_b200.h declares; compute entry
points fail loudly without a device (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "spe_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spe_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from spe_b200 import _lib, build
    so = build.build_native()
    L = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    # the ctypes table covers the header
    assert set(names) <= set(_lib.EXPORTS) | {"spe_last_error"}, sorted(set(names) - set(_lib.EXPORTS))
    assert _lib.lib().spe_version() == 100


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "spe_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_ops_fail_loudly_without_cuda():
    from spe_b200 import ops
    x = torch.zeros(4, 8, dtype=torch.bfloat16)
    w = torch.zeros(8, 8)
    with pytest.raises(RuntimeError):
        ops.linear(x, w)
    from spe_b200.models.conditional_detr import ConditionalDETR_Refine  # noqa: F401  (imports fine on CPU)
    from spe_b200 import factory
    from oracle import spe_oracle as O
    model = factory.build_detector(O.tiny_config(), "cpu")
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 32, 32))
