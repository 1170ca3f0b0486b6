"""CPU: libb200whisper.so loads and exports every symbol include/b200w.h declares; the ctypes
table covers the same set (no compute calls here: there is no GPU in the dev container)."""
import ctypes
import os
import re

from realtime_whisper_asr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "b200w.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(b200w_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in b200w.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, set(names) ^ set(_lib.SIGNATURES)


def test_load_and_version():
    lib = _lib.load()
    assert lib.b200w_abi_version() == 3
    assert lib.b200w_launch_count() >= 0
    assert lib.b200w_last_error() is not None


def test_errors_are_reported_not_swallowed():
    lib = _lib.load()
    rc = lib.b200w_engine_create(None, None)
    assert rc == -1 and b"null" in lib.b200w_last_error()


def test_sass_is_blackwell_native():
    """The built library carries sm_100a SASS with tcgen05 MMA, TMEM loads and TMA loads."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass  # no legacy mma.sync tensor path
