"""Importable alias of the ``realtime-whisper-asr_b200/`` package directory.

The package directory carries the reference's name (with hyphens, which Python cannot import);
this shim makes ``import realtime_whisper_asr_b200`` resolve to it.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "realtime-whisper-asr_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
