"""Per-source-line instruction counts and stall samples of one kernel from an ncu report (needs -lineinfo builds).

    python scripts/ncu_by_line.py <report.ncu-rep> <object-or-.so with the kernel> <kernel name substring> [top N]

ncu's CSV source page is SASS only; the line table comes from `nvdisasm -g` of the same cubin, matched by instruction
offset.  Prints the lines with the most executed warp instructions and the most stall samples, plus totals per line range
given as start:end:name arguments (phases).
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile


def line_table(obj, kernel):
    d = tempfile.mkdtemp()
    import os
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    import glob
    tab = {}
    for cubin in glob.glob(d + "/*.cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        cur, on, inl = None, False, None
        for ln in out.splitlines():
            if ln.startswith("//--------------------- .text."):
                on = kernel in ln
                continue
            if not on:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                # inlined code carries "inlined at ..., line N": attribute to the outermost (kernel) line
                m2 = re.findall(r'inlined at "[^"]+", line (\d+)', m.group(3))
                cur = int(m2[-1]) if m2 else int(m.group(2))
                continue
            m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
            if m and cur is not None:
                tab[int(m.group(1), 16)] = cur
    return tab


def main():
    rep, obj, kernel = sys.argv[1:4]
    phases = [a for a in sys.argv[4:] if ":" in a]
    top = next((int(a) for a in sys.argv[4:] if a.isdigit()), 25)
    tab = line_table(obj, kernel)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    launch = next((int(a.split("=")[1]) for a in sys.argv[4:] if a.startswith("launch=")), 0)
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]      # one section per profiled launch
    if starts:
        lo = starts[min(launch, len(starts) - 1)]
        hi_ = starts[launch + 1] if launch + 1 < len(starts) else len(rows)
        rows = rows[lo:hi_]
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    base = None
    inst, samp = collections.Counter(), collections.Counter()
    stall = collections.defaultdict(collections.Counter)
    scols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in rows[hi + 1:]:
        if len(r) < 40 or not r[0].startswith("0x"):
            continue
        a = int(r[0], 16)
        base = a if base is None else base
        line = tab.get(a - base, -1)
        inst[line] += int(r[ix["Instructions Executed"]] or 0)
        samp[line] += int(r[ix["# Samples"]] or 0)
        for h in scols:
            stall[line][h] += int(r[ix[h]] or 0)
    ti, ts = sum(inst.values()), sum(samp.values())
    print(f"total warp instructions {ti}, stall samples {ts}")
    srcfile = next((a.split("=")[1] for a in sys.argv[4:] if a.startswith("src=")), "eventclip_b200/csrc/event2img.cu")
    src = open(srcfile).read().splitlines()
    print("\n-- lines by executed warp instructions")
    for line, n in inst.most_common(top):
        t = src[line - 1].strip()[:90] if 0 < line <= len(src) else ""
        top_st = ", ".join(f"{k[6:]}={v}" for k, v in stall[line].most_common(3))
        print(f"{line:6d} {100 * n / ti:5.1f}% inst {100 * samp[line] / max(ts, 1):5.1f}% samples  [{top_st}]  {t}")
    print("\n-- lines by stall samples")
    for line, n in samp.most_common(top):
        t = src[line - 1].strip()[:90] if 0 < line <= len(src) else ""
        top_st = ", ".join(f"{k[6:]}={v}" for k, v in stall[line].most_common(3))
        print(f"{line:6d} {100 * inst[line] / ti:5.1f}% inst {100 * n / max(ts, 1):5.1f}% samples  [{top_st}]  {t}")
    if phases:
        print("\n-- phases")
        for ph in phases:
            a, b, name = ph.split(":", 2)
            a, b = int(a), int(b)
            pi = sum(n for l, n in inst.items() if a <= l <= b)
            ps = sum(n for l, n in samp.items() if a <= l <= b)
            agg = collections.Counter()
            for l in stall:
                if a <= l <= b:
                    agg.update(stall[l])
            top_st = ", ".join(f"{k[6:]}={100 * v / max(ps, 1):.0f}%" for k, v in agg.most_common(4))
            print(f"{name:28s} lines {a}-{b}: {100 * pi / ti:5.1f}% inst ({pi}), {100 * ps / max(ts, 1):5.1f}% samples  [{top_st}]")


if __name__ == "__main__":
    main()
