"""Test harness: contract-faithful stand-ins for the un-vendored upstream modules the reference
imports (`whisper_online`, `silero_vad_iterator`, a Silero VAD model), restated from SURVEY.md
Appendix A.1.  Test infrastructure only; injected through sys.modules so that the UNCHANGED
reference files import and run on top of the B200 backend."""
import os
import sys


def install(reference_dir="/root/reference"):
    """Register the stand-ins and make the reference importable.  Returns True when the reference
    tree is present (dev container), False otherwise (GPU box)."""
    from . import silero_vad_iterator, whisper_online
    sys.modules.setdefault("whisper_online", whisper_online)
    sys.modules.setdefault("silero_vad_iterator", silero_vad_iterator)
    if os.path.isdir(reference_dir):
        if reference_dir not in sys.path:
            sys.path.insert(0, reference_dir)
        return True
    return False
