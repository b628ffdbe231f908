"""Turns ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_v2.csv profiles/r01_launches_step.txt
    python scripts/summarize_ncu.py raw gpurun_out/prof_gemm_v2.ncu-rep profiles/r01_gemm_ncu.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpc__cycles_elapsed.avg.per_second", "launch__shared_mem_per_block_dynamic"]


def launches(src, dst):
    lines = [l for l in open(src) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {h: i for i, h in enumerate(hdr)}
    seq = []
    for row in r:
        name = row[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        v = float(row[ix["Metric Value"]].replace(",", ""))
        u = row[ix["Metric Unit"]]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        seq.append((name, v))
    starts = [i for i, (n, _) in enumerate(seq) if "event2img" in n]
    step = seq[starts[-1]:] if starts else seq
    agg = collections.OrderedDict()
    for n, v in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in step)
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: last bench step ({len(step)} launches, "
                f"{tot:.1f} us of kernel time; cold-cache serialised launches: compare shares, not absolutes)\n")
        f.write(f"{'share':>7} {'total_us':>10} {'launches':>8}  kernel\n")
        for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100 * v / tot:6.1f}% {v:10.1f} {c:8d}  {n}\n")
        f.write("\n# launch order of the first transformer block\n")
        for n, v in step[:12]:
            f.write(f"{v:10.1f} us  {n}\n")


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if any(h == k or h.endswith("." + k) for k in KEYS)]
    stall = [i for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    kn = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none, summary of {src} (per launch)\n")
        for r in rows[2:]:
            f.write(f"\n== {r[kn]}\n")
            for i in cols:
                f.write(f"  {hdr[i]} [{units[i]}] = {r[i]}\n")
            st = sorted([(float(r[i]) if r[i] else 0.0, hdr[i].split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""))
                         for i in stall], reverse=True)[:6]
            f.write("  top stall reasons (warps per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in st) + "\n")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
