"""The C-ABI shared library loads and exports every symbol include/navmodel_b200.h declares (no
compute calls: this runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "navmodel_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from navigation_model_b200 import _native
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_native.library_path())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_native.SIGNATURES), set(names) ^ set(_native.SIGNATURES)


def test_version_error_and_no_silent_cpu_fallback():
    from navigation_model_b200 import _native
    lib = _native.lib()
    assert lib.nm_version() == 2
    assert lib.nm_mouse_num_actions(None) < 0
    assert b"NULL" in lib.nm_last_error()
    assert _native.STEP_REC_DTYPE.itemsize == 32
    if _native.device_count() == 0:
        import pytest
        with pytest.raises(RuntimeError):
            _native.require_device()


def test_sass_is_blackwell_native():
    """The built library carries sm_100a code with the TMA bulk copy the fit kernel stages streams with."""
    import shutil
    import subprocess
    from navigation_model_b200 import _native
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _native.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass and "DFMA" in sass
