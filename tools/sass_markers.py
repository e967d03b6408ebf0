"""SASS evidence for the built library: per kernel, how many of the instructions that prove the sm_100a features in use
(cuobjdump -sass of mojo_simdjson_b200/libsimdjson_b200.so).  Writes profiles/sass_markers.txt, stamped with the git SHA.

    UBLKCP   cp.async.bulk (bulk global -> shared copies)          SYNCS    mbarrier operations
    LOP3     three-input logic (the bit-plane classifier)           PRMT     byte permute (4x4 byte transpose)
    VOTE     ballots                                                REDUX    warp reductions
    FLO      find-leading-one (index extraction)                    POPC     population count
    STS/LDS  shared memory, STG.E.128 / LDG.E.128 vector global access, ATOMG / RED atomics, ACQBULK / griddepcontrol (PDL)
usage: python tools/sass_markers.py"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mojo_simdjson_b200", "libsimdjson_b200.so")
MARKERS = ["UBLKCP", "SYNCS", "LOP3", "PRMT", "VOTE", "REDUX", "FLO", "POPC", "STS", "LDS", "STG.E.128", "LDG.E.128", "ATOMG", "RED", "ACQBULK", "MEMBAR", "NANOSLEEP", "CCTL"]


def main():
    sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    dirty = subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "--", "mojo_simdjson_b200/csrc", "include"], capture_output=True, text=True).stdout.strip()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0].replace("void ", "")
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            per[cur]["total"] += 1
            for k in MARKERS:
                if op == k or op.startswith(k + ".") or (("." in k) and op.startswith(k)):
                    per[cur][k] += 1
    out = os.path.join(ROOT, "profiles", "sass_markers.txt")
    with open(out, "w") as f:
        f.write(f"# cuobjdump -sass mojo_simdjson_b200/libsimdjson_b200.so | per-kernel instruction counts (static)\n")
        f.write(f"# git {sha}{' + uncommitted changes in csrc/' if dirty else ''}; cubin arch: {', '.join(arch)}; nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3\n")
        f.write("%-52s %6s " % ("kernel", "total") + " ".join("%9s" % k for k in MARKERS) + "\n")
        for name, c in per.items():
            f.write("%-52s %6d " % (name[:52], c["total"]) + " ".join("%9d" % c[k] for k in MARKERS) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    sys.exit(main())
