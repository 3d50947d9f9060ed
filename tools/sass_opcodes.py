#!/usr/bin/env python
"""Per-kernel SASS opcode counts of libtrxlppo.so (cuobjdump -sass; runs without a GPU) -> profiles/<round>_sass_opcodes.txt.

    python tools/sass_opcodes.py --round r2
"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "LDGSTS", "SYNCS", "HMMA", "FFMA"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r2")
    args = ap.parse_args()
    lib = os.path.join(ROOT, "episodic-transformer-memory-ppo_b200", "libtrxlppo.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts, name = collections.OrderedDict(), None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1).split(".")[0]
            if op in OPS:
                counts[name][op] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    out = os.path.join(ROOT, "profiles", "%s_sass_opcodes.txt" % args.round)
    with open(out, "w") as f:
        f.write("# SASS opcode counts per kernel of libtrxlppo.so (cuobjdump -sass, sm_100a), final build of the round (tools/sass_opcodes.py)\n")
        f.write("# tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, TMA cp.async.bulk.tensor -> UTMALDG, 1-D cp.async.bulk -> UBLKCP,\n")
        f.write("# prefetch.tensormap -> UTMAPF, cp.async -> LDGSTS, mbarrier -> SYNCS.  Kernels without any of these are omitted except for a total line.\n")
        f.write("%-92s" % "kernel" + "".join("%9s" % o for o in OPS) + "\n")
        shown = 0
        for (mangled, c), dem in zip(counts.items(), demangled):
            if not any(c[o] for o in OPS[:11]):
                continue
            short = re.sub(r"\(anonymous namespace\)::", "", dem)
            short = re.sub(r"\(.*$", "", short)
            f.write("%-92s" % short[:92] + "".join("%9d" % c[o] for o in OPS) + "\n")
            shown += 1
        f.write("# %d kernels in the library, %d listed\n" % (len(counts), shown))
    print("wrote", out)


if __name__ == "__main__":
    main()
