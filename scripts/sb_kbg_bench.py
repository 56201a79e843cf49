"""Small query blocks (B <= 16, smallbatch kernel): K-slabs per TMA box sweep.  Checks results against the scan."""
import os, sys, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
M, D, K = 1_000_000, 768, 10
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
for dt in (torch.float32, torch.bfloat16):
    rows = torch.randn(M, D, device=dev, generator=g).to(dt)
    inv = ops.row_inv_norms(rows)
    for B in (8, 16):
        q = torch.randn(B, D, device=dev, generator=g)
        ref_i, ref_s = ops.scan_topk(rows, q, K, inv, None)
        for kbg in (1, 2, 4, 1, 2, 4):
            os.environ["AURA_SB_KBG"] = str(kbg)
            ts = []
            for i in range(8):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); idx, sc, unc = ops.batch_topk(rows, q, K, inv, None); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ok = torch.equal(idx, ref_i) if int(unc.sum()) == 0 else "uncertified:%d" % int(unc.sum())
            print(f"{str(dt)[6:]} B={B} kbg={kbg} us: " + " ".join(f"{t:.0f}" for t in ts[2:]), "idx_equal_scan", ok, flush=True)
