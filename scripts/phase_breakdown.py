"""Per-phase instruction / stall-sample shares of one kernel launch from an ncu report (phases split at barriers)."""
import collections
import csv
import subprocess
import sys

rep, which = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
allrows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
rows = allrows[starts[which]:starts[which + 1]]
print("=====", rows[0][1][:70])
hdr = rows[1]
ix, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
data = [(int(r[ix]) if r[ix].isdigit() else 0, int(r[ismp]) if r[ismp].isdigit() else 0, r[isrc].strip()) for r in rows[2:] if len(r) > ix]
tot, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print("total warp-instr", tot, "samples", ts, "sass lines", len(data))
marks = [i for i, d in enumerate(data) if ("BAR.SYNC" in d[2] or "UCGABAR_WAIT" in d[2])]
prev = 0
for m in marks + [len(data) - 1]:
    seg = data[prev:m + 1]
    c, s = sum(d[0] for d in seg), sum(d[1] for d in seg)
    if c > tot * 0.01 or s > ts * 0.015:
        ops = collections.Counter()
        smp = collections.Counter()
        for d in seg:
            t = d[2].split()
            if not t:
                continue
            k = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            ops[k] += d[0]
            smp[k] += d[1]
        print(f"[{prev:5d},{m:5d}] instr {100*c/tot:5.1f}% samples {100*s/ts:5.1f}% | instr:",
              [(k, round(100 * v / tot, 1)) for k, v in ops.most_common(5)], "| samples:",
              [(k, round(100 * v / ts, 1)) for k, v in smp.most_common(4)])
    prev = m + 1
