#!/bin/bash
# Round 2, GPU call T: debug of integration_frame x transposed kernel, then the tests that failed in call S.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T=${TAG:-r02t}
python -c "import nyx_b200.abi as a; a.load_library()" || { echo "libnyxb.so missing or stale"; exit 9; }
timeout 300 python scripts/debug/dbg_frame_tx.py 2>&1 | tee gpurun_out/${T}_dbg_frame_tx.log | tail -40
timeout 900 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_frames_fields.py tests/test_gpu_tx.py tests/test_trajectory.py tests/test_gpu_fullsize.py > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${T}_pytest.log
