"""Text summary of one `ncu --set full --import-source on` capture (first kernel in the report).

usage: python profiles/tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/r01_ncu_x.txt
Prints duration, tensor-pipe %, DRAM / L2->SM bytes, issue utilisation, registers, the aggregate warp-stall
sampling breakdown and the hottest SASS lines.  Needs the `ncu` CLI (reads the report, no GPU)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("report:", rep)
    print("kernel:", d.get("Kernel Name", ("?", ""))[0][:160])
    for k in KEYS:
        if k in d:
            print(f"  {k:72s} {d[k][0]:>16s} {d[k][1]}")
    stalls = {h: float(v) for h, v in ((h, d[h][0]) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio"))}
    print("warp stalls per issue-active cycle (top):")
    for h, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:6.2f}")
    src = page(rep, "source")
    if len(src) > 2:
        sh = src[1]
        rows = [r for r in src[2:] if len(r) == len(sh)]
        ix = {h: i for i, h in enumerate(sh)}
        if "# Samples" in ix:
            tot = sum(int(r[ix["# Samples"]] or 0) for r in rows)
            names = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
            print(f"hottest SASS lines (of {tot} samples):")
            for r in sorted(rows, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]:
                n = int(r[ix["# Samples"]] or 0)
                top = max(names, key=lambda h: int(r[ix[h]] or 0))
                print(f"  {100.0 * n / max(tot, 1):5.1f}%  {top:22s} {r[ix['Source']].strip()[:80]}")


if __name__ == "__main__":
    main()
