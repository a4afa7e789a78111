#!/usr/bin/env python
"""Summarise `ncu --set full` captures (raw-page CSVs written by `ncu -i x.ncu-rep --page raw --csv`) of the hot kernels
into profiles/<round>_ncu/traffic.json, which bench.py reads for `roofline.traffic`:

    python tools/ncu_summary.py gpurun_out/r2_ncu profiles/r2_ncu

For every <target>.raw.csv: kernel name, duration, dram bytes read + written, tensor-pipe / issue activity, registers.
Per-layer aggregates: the four decode GEMM launches (gemm_qkv + gemm_o + gemm_gate_up + gemm_down) and attention."""
import csv
import json
import os
import shutil
import sys

KEYS = {
    "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum", "duration_ns": "gpu__time_duration.sum",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "grid": "launch__grid_size",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9,
        "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}


def parse(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = {"kernel": r[names.index("Kernel Name")]}
        for k, m in KEYS.items():
            col = next((i for i, n in enumerate(names) if n.endswith(m)), None)
            if col is None:
                continue
            try:
                d[k] = float(r[col].replace(",", "")) * UNIT.get(units[col], 1.0)
            except ValueError:
                pass
        out.append(d)
    return out


def main(src, dst):
    os.makedirs(dst, exist_ok=True)
    summary = {}
    for f in sorted(os.listdir(src)):
        if not f.endswith(".raw.csv"):
            continue
        tgt = f[:-len(".raw.csv")]
        try:
            recs = parse(os.path.join(src, f))
        except Exception as e:  # noqa: BLE001
            print("skip", f, e)
            continue
        if not recs:
            continue
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
        r = recs[-1]   # the measured (last) launch; earlier ones are warm-ups
        r["dram_bytes"] = r.get("dram_read", 0.0) + r.get("dram_write", 0.0)
        summary[tgt] = r
    agg = {}
    g = [summary.get(f"gemm_{n}") for n in ("qkv", "o", "gate_up", "down")]
    if all(g):
        agg["w4a8_gemm(decode,4 launches/layer)"] = sum(x["dram_bytes"] for x in g)
    if "attention" in summary:
        agg["kv4_decode_attention"] = summary["attention"]["dram_bytes"]
    json.dump({"source": "ncu --set full --clock-control none, one cold launch per kernel (tools/ncu_targets.py)",
               "per_layer_dram_bytes": agg, "kernels": summary}, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
    for k, v in summary.items():
        print(f"{k:18s} {v['kernel'][:60]:60s} {v.get('duration_ns', 0) / 1e3:8.1f} us  dram {v['dram_bytes'] / 1e6:8.2f} MB  "
              f"tensor {v.get('tensor_pct', 0):5.1f}%  issue {v.get('issue_pct', 0):5.1f}%  regs {int(v.get('regs', 0))}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
