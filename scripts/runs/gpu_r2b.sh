#!/bin/bash
# round 2, call b: full -m gpu suite after the ABI / hygiene / exact-coarse changes + new scale tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --durations=12 2>&1 | tail -40 | tee gpurun_out/r2b_pytest.log
