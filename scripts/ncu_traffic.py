"""Read `ncu --set full` reports (gpurun_out/*.ncu-rep, on the CPU box) and write, per captured kernel launch, the numbers
the bench and DESIGN.md cite: duration, DRAM bytes read + written, DRAM / tensor-pipe / L2 utilisation, registers.
usage: python scripts/ncu_traffic.py key=report.ncu-rep[:launch_index[:kernel substring]] ... ; updates
profiles/ncu_traffic.json (bench.py's roofline.traffic source) and prints a table for profiles/r02_ncu_summary.txt."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pct",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct", "launch__registers_per_thread": "regs",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "smsp__inst_executed.sum": "warp_insts"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    res = []
    for row in r[2:]:
        d = {"kernel": row[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in WANT and row[i] not in ("", "n/a"):
                d[WANT[h]] = float(row[i].replace(",", "")) * SCALE.get(units[i], 1.0)
        res.append(d)
    return res


def main():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    for arg in sys.argv[1:]:
        key, spec = arg.split("=", 1)
        parts = spec.split(":")
        rep, idx = parts[0], int(parts[1]) if len(parts) > 1 and parts[1] else 0
        sub = parts[2] if len(parts) > 2 else ""
        rows = [r for r in rows_of(rep) if sub in r["kernel"]]
        r = rows[idx]
        table[key] = {"bytes": r.get("dram_read", 0.0) + r.get("dram_write", 0.0),
                      "source": f"profiles/{os.path.basename(rep).replace('.ncu-rep', '_raw.csv')} ({r['kernel'][:60]}, launch {idx}: "
                                f"dram__bytes_read.sum + dram__bytes_write.sum)",
                      "duration_ms": r.get("duration"), "dram_pct": r.get("dram_pct"), "tensor_pct": r.get("tensor_pct"),
                      "l2_hit_pct": r.get("l2_hit_pct"), "regs": r.get("regs")}
        print(f"{key:22s} {r['kernel'][:48]:48s} {r.get('duration', 0):9.3f} ms  DRAM {r.get('dram_read', 0) / 1e9:8.3f} + "
              f"{r.get('dram_write', 0) / 1e9:6.3f} GB ({r.get('dram_pct', 0):5.1f} %)  tensor {r.get('tensor_pct', 0):5.1f} %  "
              f"L2 hit {r.get('l2_hit_pct', 0):5.1f} %  regs {int(r.get('regs', 0))}")
    json.dump(table, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
