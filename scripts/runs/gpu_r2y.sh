#!/bin/bash
# full GPU suite + smoke + default bench line after the dense-kernel changes (two warpgroups, sampled start threshold)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2y_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench_err.log; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2y_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "ms_per_step_reps", "uncertified_queries_rerun")})
print("e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["traffic"], "clocks", d["clocks"])
print("c3", d["c3_allpairs"]["ms"], d["c3_allpairs"]["roofline"]["frac"])
c4 = d["c4_ivf"]; print("c4", {m: c4[m]["ms_per_batch"] for m in ("gather", "list_major", "list_major_bf16")}, c4["list_major_bf16_same_result"], c4["build_ms"], c4["online_writes"])
c5 = d["c5_shard"]; print("c5", c5["relaxed"]["ms_per_batch"], c5["strict"]["ms_per_batch"], c5["recall_at_10_vs_exact"], c5["roofline"]["traffic"])
PY
