"""CPU: the C-ABI library loads and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

import distill_bev_b200
from distill_bev_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names.update(re.findall(r"\b(dbev_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exists_in_tree():
    assert os.path.exists(distill_bev_b200.library_path()), "run python distill-bev_b200/build.py"
    assert distill_bev_b200.library_path().startswith(ROOT)


def test_every_declared_symbol_is_exported_and_bound():
    declared = _declared_symbols()
    assert len(declared) >= 10
    lib = ctypes.CDLL(distill_bev_b200.library_path())
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export %s" % name
    assert declared == set(_lib.SIGNATURES), (
        "ctypes table out of sync with the header: %s" % (declared ^ set(_lib.SIGNATURES)))


def test_abi_version_and_arch():
    lib = _lib.load()
    assert lib.dbev_abi_version() == 1
    assert lib.dbev_build_arch() == b"sm_100a"


def test_library_has_only_sm100a_code():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "-lelf", distill_bev_b200.library_path()],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_cpu_tensor_is_rejected_loudly():
    import pytest
    import torch
    with pytest.raises(RuntimeError, match="no CPU path"):
        distill_bev_b200.bev_pool(torch.zeros(4, 8), torch.zeros(4, 4, dtype=torch.long), 1, 1, 2, 2)
