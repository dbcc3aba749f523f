"""Per-layer table of one Oobleck decode from an ncu launch list.

usage: python profiles/tools/decoder_layer_table.py gpurun_out/launches_dec.csv [L_latent] [first_id]

The CSV comes from
  ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,\
dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled \
      -k regex:satb -s <skip> -c <count> --csv --log-file <csv> python tests/prof_step.py oobleck 1024
(durations under ncu are cold-cache and serialised: use them for shares, not as bench values).
Algorithmic bytes follow DESIGN.md: 16-bit activated copies, fp32 raw skip stream.
"""
import collections
import csv
import sys


def layers(L, fused128=True):
    """(name, flops, algorithmic bytes) in launch order for the SA-Open-1.0 decoder.  With fused128 the
    128- and 256-channel ResidualUnits are one launch each (resunit[256]_tcgen05_2cta_kernel)."""
    out = [("ncl->nlc16", 0, 64 * L * 6)]
    out.append(("conv_in k7 64->2048", 2 * L * 64 * 2048 * 7, L * (64 * 2 + 2048 * 6)))
    cin = 2048
    for s, cout in zip((8, 8, 4, 4, 2), (1024, 512, 256, 128, 128)):
        lo = L * s
        out.append((f"convT s{s} {cin}->{cout}", 2 * L * cin * cout * 2 * s, L * cin * 2 + lo * cout * 6))
        for d in (1, 3, 9):
            if fused128 and cout in (128, 256):
                out.append((f"  resunit d{d} {cout} (fused)", 2 * lo * cout * cout * 8, lo * cout * (2 + 4 + 4 + 2)))
                continue
            out.append((f"  conv7 d{d} {cout}", 2 * lo * cout * cout * 7, lo * cout * 4))
            out.append((f"  conv1+skip {cout}", 2 * lo * cout * cout, lo * cout * (2 + 4 + 4 + 2)))
        cin, L = cout, lo
    out.append(("conv_out k7 128->2", 2 * L * 128 * 2 * 7, L * (128 * 2 + 2 * 4)))
    return out


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
    spec = layers(L, fused128="--unfused" not in sys.argv)
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
