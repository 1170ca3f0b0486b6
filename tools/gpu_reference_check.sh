#!/bin/bash
# Run tests/test_gpu_reference_files.py on a GPU box: the reference tree does not exist there, so a TRANSIENT copy of
# its .py files is placed in the git-ignored tests/_reference_copy/ (it travels with the gpurun snapshot) and removed
# afterwards.  Run from the dev container:  bash tools/gpu_reference_check.sh
set -e
cd "$(dirname "$0")/.."
mkdir -p tests/_reference_copy
cp /root/reference/asr_components.py /root/reference/enhanced_asr_processor.py /root/reference/audio_deduplicator.py tests/_reference_copy/
trap 'rm -rf tests/_reference_copy' EXIT
/usr/local/graft/bin/gpurun --timeout 600 -- 'python -m pytest tests/test_gpu_reference_files.py -m gpu -q -s 2>&1 | tail -8 | cut -c1-250'
