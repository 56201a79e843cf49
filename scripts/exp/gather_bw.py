"""Experiment: random 4 KB row gathers vs contiguous rows on a 41 GB bank (address-translation / DRAM-page cost)."""
import torch
dev = torch.device("cuda:0")
for M in (1_000_000, 10_000_000):
    bank = torch.empty(M, 1024, device=dev)
    bank.normal_()
    n = 1 << 19
    for name, idx in (("random", torch.randint(0, M, (n,), device=dev)), ("contiguous", torch.arange(n, device=dev)),
                      ("random-sorted", torch.randint(0, M, (n,), device=dev).sort().values)):
        out = torch.empty(n, 1024, device=dev)
        for _ in range(2):
            torch.index_select(bank, 0, idx, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            torch.index_select(bank, 0, idx, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"M={M} {name:14s}: {ms:.3f} ms  read {n*4096/ms/1e6:.0f} GB/s (+ same written)")
    del bank
