"""Cut ONE denoise step out of an ncu launch list of `bench.py` and sum it by kernel class.

usage: python profiles/tools/step_from_launches.py gpurun_out/r02_launches_step.csv [step_index]

The list (ncu --metrics gpu__time_duration.sum --csv) holds weight preparation, conditioning preparation and several
steps; a step starts at the `fourier_kernel` launch (first kernel of satb_dit_forward) and ends before the next one.
Durations under ncu are serialised and cold-cache: use the SHARES, not the absolute times."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rows = list(csv.DictReader([l for l in open(path, errors="ignore") if not l.startswith("==")]))
    launches = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        us = v / 1e3 if u.startswith("ns") else (v if u.startswith("us") else v * 1e3)
        launches.append((r["Kernel Name"], us))
    starts = [i for i, (n, _) in enumerate(launches) if "fourier_kernel" in n]
    if len(starts) <= which + 1:
        which = max(0, len(starts) - 2)
    a, b = starts[which], starts[which + 1]
    step = launches[a:b]

    def cls(n):
        for key, name in (("EpiSwiglu", "FF-in GEMM (bias + SwiGLU)"), ("EpiQkvRope", "QKV GEMM (+RoPE)"),
                          ("EpiResidual", "residual GEMMs (attn out, cross out, FF-out)"), ("EpiStore16", "cross q GEMM"),
                          ("attn_tc_kernel", "attention (self + cross)"), ("layernorm_kernel", "LayerNorm"),
                          ("EpiStore32", "project in / out GEMM"), ("sampler_update", "sampler update")):
            if key in n:
                return name
        return "other (" + re.sub(r"<.*", "", n.split("(")[0]).split("::")[-1][:40] + ")"

    agg = collections.OrderedDict()
    for n, us in step:
        d = agg.setdefault(cls(n), [0, 0.0])
        d[0] += 1
        d[1] += us
    tot = sum(us for _, us in step)
    print(f"step {which}: {len(step)} launches, {tot / 1e3:.2f} ms of kernel time under ncu")
    for k, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:48s} {cnt:4d} launches {us / 1e3:8.3f} ms  {100 * us / tot:5.1f} %")


if __name__ == "__main__":
    main()
