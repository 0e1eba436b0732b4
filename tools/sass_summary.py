#!/usr/bin/env python
"""tools/sass_summary.py — opcode histogram per kernel of htool_b200/lib/libhtool_b200.so (cuobjdump -sass), the evidence
that the kernels are sm_100a-native: UBLKCP (1-D bulk TMA copies), SYNCS.* (mbarriers), DMMA.8x8x4 (FP64 tensor cores),
LDGSTS (cp.async), LDS.128, no library kernels. Writes profiles/<round>_sass_summary.txt.

  python tools/sass_summary.py [output.txt]
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "htool_b200", "lib", "libhtool_b200.so")
KEY = ["UBLKCP", "UTMALDG", "SYNCS", "DMMA", "LDGSTS", "LDS", "STS", "LDG", "STG", "DFMA", "DADD", "DMUL", "SHFL", "BAR", "ATOM", "RED", "NANOSLEEP", "ACQBULK", "UTCMMA", "LDTM"]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "profiles", "r02_sass_summary.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels, cur, arch = collections.OrderedDict(), None, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch = m.group(1)
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    names = list(kernels)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
    for n, d in zip(names, dem):
        demangle[n] = re.sub(r"htb::\(anonymous namespace\)::|htb::", "", d)
    total = collections.Counter()
    with open(out_path, "w") as f:
        f.write(f"# cuobjdump -sass {os.path.relpath(LIB, REPO)} — arch {arch}; static instruction counts per kernel (opcode families)\n")
        for n, c in kernels.items():
            fam = collections.Counter()
            for op, k in c.items():
                fam[op.split(".")[0]] += k
                total[op] += k
            n_inst = sum(c.values())
            keys = {k: fam[k] for k in KEY if fam.get(k)}
            dm = sum(v for k, v in c.items() if k.startswith("DMMA"))
            f.write(f"\n{demangle.get(n, n)}\n  instructions {n_inst}; " + ", ".join(f"{k} {v}" for k, v in keys.items()) + "\n")
            detail = {k: v for k, v in c.items() if k.startswith(("DMMA", "UBLKCP", "SYNCS", "LDGSTS", "LDS.128", "LDS.64"))}
            if detail:
                f.write("  " + ", ".join(f"{k} {v}" for k, v in sorted(detail.items())) + "\n")
        f.write("\n# whole library\n")
        for k in ("UBLKCP", "UTMALDG", "UTCMMA", "LDTM"):
            f.write(f"{k}: {sum(v for op, v in total.items() if op.startswith(k))}\n")
        f.write(f"SYNCS.*: {sum(v for op, v in total.items() if op.startswith('SYNCS'))}\n")
        f.write(f"DMMA.8x8x4: {sum(v for op, v in total.items() if op.startswith('DMMA'))}\n")
        f.write(f"LDGSTS: {sum(v for op, v in total.items() if op.startswith('LDGSTS'))}\n")
        f.write(f"LDS.128: {sum(v for op, v in total.items() if op.startswith('LDS') and '.128' in op)}\n")
        f.write("(no UTCMMA / LDTM: there is no FP64 tcgen05; no UTMALDG: the streams are 1-D, tile TMA is not needed)\n")
    print(out_path)


if __name__ == "__main__":
    main()
