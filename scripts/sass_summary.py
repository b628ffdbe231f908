"""Per-kernel SASS opcode evidence from the built library (no GPU needed):

    python scripts/sass_summary.py [eventclip_b200/libeventclip_b200.so] > profiles/r02_sass_opcodes.txt

Counts, for every kernel of the .so, the mnemonics that prove which hardware path it uses (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor loads / stores, UTCBAR = tcgen05.commit,
IMMA / HMMA = legacy mma.sync (int8 / half), ATOMS = shared-memory atomics, ATOM / RED = global atomics, SYNCS = mbarrier ops,
UCGABAR = cluster barriers, MATCH = match.any / all."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "eventclip_b200/libeventclip_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS",
         "UCGABAR", "HMMA", "IMMA", "ATOMS", "ATOMG", "ATOM", "RED", "REDUX", "MATCH", "MUFU.EX2", "LDGSTS", "F2FP", "DFMA", "DMUL"]
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for ln in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[kern][w] += 1
                break
print(f"# cuobjdump -sass {lib}: static instruction counts per kernel, watched mnemonics only")
print("# tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG, mma.sync -> HMMA/IMMA, shared atomics -> ATOMS")
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
for (k, c), name in zip(counts.items(), names):
    name = re.sub(r"\(anonymous namespace\)::", "", name).split("(E2I")[0].split("(ec")[0]
    name = re.sub(r"^void ", "", name)[:84]
    ops = ", ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{name:84s} {total[k]:6d} instr  {ops}")
