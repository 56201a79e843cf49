#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x --deselect tests/test_ivf_scale_gpu.py 2>&1 | tail -12 | tee gpurun_out/r2l_pytest.log
timeout 300 python bench.py --steps 10 --no-cpu-baseline --legs c3 2>gpurun_out/r2l_err.log > gpurun_out/r2l_bench.json; tail -2 gpurun_out/r2l_err.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2l_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["uncertified_queries_rerun"], d["roofline"]["frac"])
print(d["c3_allpairs"]["ms"], d["c3_allpairs"]["roofline"]["frac"], d["c3_allpairs"]["top32_overlap_vs_fp32_sample"])
PY
