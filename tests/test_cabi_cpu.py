"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/movii_b200.h declares, and fails
loudly (no fallback) when there is no B200."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "movii_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(mv):
    lib = mv.lib()
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), "symbol %s declared in include/movii_b200.h is not exported" % name
    assert sorted(mv.EXPORTED_SYMBOLS) == declared


def test_version_and_error_channel(mv):
    lib = mv.lib()
    assert lib.mv_version() >= 100
    assert isinstance(lib.mv_last_error(), bytes)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_fails_loudly_without_gpu(mv):
    with pytest.raises(RuntimeError):
        mv.device_check()
    x = torch.zeros(4, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        mv.ln_modulate(x, torch.zeros(4, 8, dtype=torch.bfloat16))
