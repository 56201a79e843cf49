"""The reference arm of bench.py (`--impl reference`: the oracle port timed on host cores) must emit the contract's JSON
line.  Runs on a shrunken bank so that the CPU suite stays fast; the full-size run is the driver's."""
import argparse
import importlib.util
import io
import json
import os
from contextlib import redirect_stdout


def test_reference_arm_emits_the_contract_line(monkeypatch):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(bench, "N_ROWS", 4000)
    monkeypatch.delenv("RANK", raising=False)
    args = argparse.Namespace(steps=2, warmup=1, ref_queries_per_step=2, gpus=1, batch=1024)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(args)
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "queries/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 2
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"].startswith("C2") and line["config"]["batch"] == 1024
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_of_the_reference_arm_do_nothing(monkeypatch):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_under_test2", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setenv("RANK", "1")
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(argparse.Namespace(steps=1, warmup=0, ref_queries_per_step=1, gpus=2, batch=1024))
    assert buf.getvalue() == ""
