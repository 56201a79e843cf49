#!/bin/bash
# two epilogue warpgroups + prefetched column terms: parity tests, then C2 shadow kernel sustained, WG=1 vs WG=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tensorcore_gpu.py tests/test_scan_gpu.py tests/test_round2_gpu.py -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2v_pytest.log
for cfg in "1 32" "2 32" "2 24" "2 48"; do set -- $cfg; echo -n "WG=$1 L=$2: "; AURA_GEMM_WG=$1 AURA_SHADOW_L=$2 timeout 300 python scripts/c2_shadow_one.py 200 2>&1 | tail -1; done | tee gpurun_out/r2v.log
