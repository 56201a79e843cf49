"""Ad-hoc timing of aura_scan_topk under the tuning knobs (not the contract bench; see bench.py).
usage: quick_scan_bench.py [one]   ('one' = a single fp32 B=1 config, for ncu)"""
import os, sys, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops

dev = torch.device("cuda:0")
n, d = 1_000_000, 768
rows32 = torch.randn(n, d, device=dev)
inv = ops.row_inv_norms(rows32)

def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, e in evs:
        a.record(); fn(); e.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(e) for a, e in evs)
    return ts[len(ts) // 2], ts[0]

def run(tag, rows, b, k, env):
    for kk in ("AURA_SCAN_RPW", "AURA_SCAN_STAGES", "AURA_SCAN_GRID", "AURA_SCAN_INTERLEAVE"):
        os.environ.pop(kk, None)
    os.environ.update({a: str(v) for a, v in env.items()})
    q = torch.randn(b, d, device=dev)
    med, mn = timeit(lambda: ops.scan_topk(rows, q, k, scale=inv))
    byt = rows.numel() * rows.element_size()
    print(f"{tag:34s} {str(rows.dtype)[6:]:8s} B={b} k={k} env={env}: median {med*1e3:7.1f} us  min {mn*1e3:7.1f} us  "
          f"{byt/med/1e6:6.0f} GB/s  qps {b/med*1e3:.0f}", flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "one":
    env = {}
    for a in sys.argv[2:]:
        kk, v = a.split("="); env[kk] = v
    run("ncu", rows32, 1, 10, env)
    sys.exit(0)

med, mn = timeit(lambda: rows32.sum())
print(f"torch.sum(bank) read-only reference: {med*1e3:.1f} us -> {rows32.numel()*4/med/1e6:.0f} GB/s")
med, mn = timeit(lambda: torch.mv(rows32, rows32[0]))
print(f"torch.mv(bank, q) (cuBLAS gemv): {med*1e3:.1f} us -> {rows32.numel()*4/med/1e6:.0f} GB/s")
run("default", rows32, 1, 10, {})
for il in (0, 1):
    for rpw, st in ((2, 4), (2, 3), (1, 8), (1, 6), (1, 4), (3, 2), (4, 2)):
        run(f"rpw={rpw} stages={st} il={il}", rows32, 1, 10, {"AURA_SCAN_RPW": rpw, "AURA_SCAN_STAGES": st, "AURA_SCAN_INTERLEAVE": il})
for gsz in (74, 111, 148):
    run(f"grid={gsz}", rows32, 1, 10, {"AURA_SCAN_GRID": gsz, "AURA_SCAN_INTERLEAVE": 1})
rows16 = rows32.to(torch.bfloat16)
for il in (0, 1):
    run(f"bf16 il={il}", rows16, 1, 10, {"AURA_SCAN_INTERLEAVE": il})
    run(f"bf16 rpw=2 st=8 il={il}", rows16, 1, 10, {"AURA_SCAN_RPW": 2, "AURA_SCAN_STAGES": 8, "AURA_SCAN_INTERLEAVE": il})
for b in (2, 4, 8):
    run("fp32 batch", rows32, b, 10, {"AURA_SCAN_INTERLEAVE": 1})
run("fp32 k=100", rows32, 1, 100, {"AURA_SCAN_INTERLEAVE": 1})
