"""Ad-hoc timing of the tcgen05 paths (aura_batch_topk / aura_allpairs_topk) under tuning knobs."""
import os, sys, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops

dev = torch.device("cuda:0")
def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, e in evs:
        a.record(); fn(); e.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(e) for a, e in evs)
    return ts[len(ts) // 2], ts[0]

def setenv(env):
    for kk in ("AURA_GEMM_STAGES", "AURA_GEMM_GROUPS", "AURA_GEMM_L"):
        os.environ.pop(kk, None)
    os.environ.update({a: str(v) for a, v in env.items()})

what = sys.argv[1] if len(sys.argv) > 1 else "all"
n, d = 1_000_000, 768
if what in ("all", "k6", "k6one"):
    rows32 = torch.randn(n, d, device=dev)
    inv = ops.row_inv_norms(rows32)
    pick = torch.randint(0, n, (1024,), device=dev)
    qall = rows32[pick] + 0.1 * torch.randn(1024, d, device=dev)
    cfgs = [(1024, {})] if what == "k6one" else [(1024, {}), (1024, {"AURA_GEMM_STAGES": 3}), (1024, {"AURA_GEMM_GROUPS": 9}), (1024, {"AURA_GEMM_GROUPS": 37}),
            (512, {}), (256, {}), (128, {}), (32, {}), (8, {}), (4, {}), (1, {})]
    for b, env in cfgs:
        setenv(env)
        q = qall[:b].contiguous()
        med, mn = timeit(lambda: ops.batch_topk(rows32, q, 10, inv))
        _, _, fl = ops.batch_topk(rows32, q, 10, inv)
        tf = 2.0 * b * n * d / med / 1e9
        print(f"K6 fp32/tf32 B={b:5d} env={env}: median {med*1e3:8.1f} us min {mn*1e3:8.1f}  {tf:7.1f} TFLOP/s  "
              f"{n*d*4/med/1e6:6.0f} GB/s(bank once)  qps {b/med*1e3:9.0f}  uncertain {int(fl.sum())}", flush=True)
    if what == "all":
        st = {}
        med, mn = timeit(lambda: ops.exact_topk_batched(rows32, qall, 10, inv, stats=st))
        print(f"K6 full path (with fallback) B=1024: median {med*1e3:.1f} us  qps {1024/med*1e3:.0f}  uncertain/call {st['uncertain']/7:.1f}")
        rows16 = rows32.to(torch.bfloat16)
        inv16 = ops.row_inv_norms(rows16)
        for b in (1024, 128, 8):
            q = qall[:b].contiguous()
            med, mn = timeit(lambda: ops.batch_topk(rows16, q, 10, inv16))
            _, _, fl = ops.batch_topk(rows16, q, 10, inv16)
            print(f"K6 bf16 B={b:5d}: median {med*1e3:8.1f} us  {2.0*b*n*d/med/1e9:7.1f} TFLOP/s  {n*d*2/med/1e6:6.0f} GB/s  qps {b/med*1e3:9.0f} uncertain {int(fl.sum())}", flush=True)
        del rows16
    del rows32
if what == "small":
    rows32 = torch.randn(n, d, device=dev)
    inv = ops.row_inv_norms(rows32)
    pick = torch.randint(0, n, (128,), device=dev)
    qall = rows32[pick] + 0.1 * torch.randn(128, d, device=dev)
    for dt, rows, iv in (("fp32", rows32, inv),):
        for b in (1, 2, 4, 8, 16, 32, 33, 64, 128):
            q = qall[:b].contiguous()
            for env in ({"AURA_SMALLBATCH": 1}, {"AURA_SMALLBATCH": 0}):
                os.environ.update({a: str(v) for a, v in env.items()})
                med, mn = timeit(lambda: ops.batch_topk(rows, q, 10, iv), iters=10)
                _, _, fl = ops.batch_topk(rows, q, 10, iv)
                print(f"{dt} B={b:4d} {env}: {med*1e3:8.1f} us  qps {b/med*1e3:9.0f}  GB/s {n*d*4/med/1e6:6.0f} uncertain {int(fl.sum())}", flush=True)
            med, mn = timeit(lambda: ops.scan_topk(rows, q, 10, iv), iters=5)
            print(f"{dt} B={b:4d} scan: {med*1e3:8.1f} us  qps {b/med*1e3:9.0f}", flush=True)
    rows16 = rows32.to(torch.bfloat16); inv16 = ops.row_inv_norms(rows16)
    os.environ["AURA_SMALLBATCH"] = "1"
    for b in (8, 32):
        q = qall[:b].contiguous()
        med, mn = timeit(lambda: ops.batch_topk(rows16, q, 10, inv16), iters=10)
        print(f"bf16 B={b:4d}: {med*1e3:8.1f} us  qps {b/med*1e3:9.0f} GB/s {n*d*2/med/1e6:6.0f}", flush=True)
if what in ("all", "k7", "k7one"):
    for nn in ((262144,) if what == "k7one" else (65536, 262144)):
        bank = torch.randn(nn, d, device=dev).to(torch.bfloat16)
        invb = ops.row_inv_norms(bank)
        for env in ([{}] if what == "k7one" else [{}, {"AURA_GEMM_STAGES": 3}]):
            setenv(env)
            med, mn = timeit(lambda: ops.allpairs_topk(bank, 32, invb), iters=3, warm=1)
            print(f"K7 all-pairs bf16 N={nn} top-32 env={env}: median {med:8.2f} ms  {2.0*nn*nn*d/med/1e9:7.1f} TFLOP/s", flush=True)
