"""Per-layer table of one Oobleck decode from an ncu launch list.

usage: python profiles/tools/decoder_layer_table.py gpurun_out/launches_dec.csv [L_latent] [first_id]

The CSV comes from
  ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,\
dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled \
      -k regex:satb -s <skip> -c <count> --csv --log-file <csv> python tests/prof_step.py oobleck 1024
(durations under ncu are cold-cache and serialised: use them for shares, not as bench values).
Algorithmic bytes: profiles/tools/decoder_bytes.py (16-bit activated copies; raw skip stream 2 B, or 4 B with --raw32).
"""
import collections
import csv
import sys


import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from decoder_bytes import layers as _layers      # the byte model bench.py uses


def layers(L, fused128=True, raw_bytes=2):
    return _layers(L, raw_bytes, fused128)


def main():
    path = sys.argv[1]
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = csv.DictReader([l for l in open(path, errors="ignore") if not l.startswith("==")])
    by = collections.OrderedDict()
    for r in rows:
        d = by.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        d["unit:" + r["Metric Name"]] = r["Metric Unit"]
    ids = [i for i in by if i >= first]
    spec = layers(L, fused128="--unfused" not in sys.argv, raw_bytes=4 if "--raw32" in sys.argv else 2)
    tot_t = tot_f = tot_b = tot_d = 0.0
    print(f"{'layer':28s} {'us':>8s} {'TF/s':>7s} {'alg GB/s':>9s} {'dram GB/s':>9s} {'tensor%':>7s}")
    for (name, fl, by_alg), i in zip(spec, ids):
        d = by[i]
        t = d["gpu__time_duration.sum"]
        u = d["unit:gpu__time_duration.sum"]
        t_us = t / 1e3 if u.startswith("ns") else (t if u.startswith("us") else t * 1e3)
        dram = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            pass
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        dram = sum(d.get(k, 0) * scale.get(d.get("unit:" + k, "byte"), 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        tens = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", float("nan"))
        print(f"{name:28s} {t_us:8.1f} {fl / t_us / 1e6:7.0f} {by_alg / t_us / 1e3:9.0f} {dram / t_us / 1e3:9.0f} {tens:7.1f}")
        tot_t += t_us; tot_f += fl; tot_b += by_alg; tot_d += dram
    print(f"{'total':28s} {tot_t:8.1f} {tot_f / tot_t / 1e6:7.0f} {tot_b / tot_t / 1e3:9.0f} {tot_d / tot_t / 1e3:9.0f}")
    print(f"flops {tot_f / 1e12:.3f} T, algorithmic bytes {tot_b / 1e9:.2f} GB, dram bytes {tot_d / 1e9:.2f} GB")


if __name__ == "__main__":
    main()
