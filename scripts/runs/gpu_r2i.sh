#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x --deselect tests/test_ivf_scale_gpu.py 2>&1 | tail -12 | tee gpurun_out/r2i_pytest.log
timeout 300 python bench.py --steps 10 --no-cpu-baseline --legs none 2>gpurun_out/r2i_err.log > gpurun_out/r2i_bench_shadow.json; tail -3 gpurun_out/r2i_err.log
timeout 300 python bench.py --steps 10 --no-cpu-baseline --legs none --no-shadow 2>gpurun_out/r2i_err2.log > gpurun_out/r2i_bench_noshadow.json; tail -2 gpurun_out/r2i_err2.log
python - <<'PY'
import json
for f in ["shadow","noshadow"]:
    d=json.load(open(f"gpurun_out/r2i_bench_{f}.json"))
    print(f, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "uncert", d["uncertified_queries_rerun"], "hit", d["top1_hit_rate"], d["roofline"]["frac"], d["clocks"])
PY
