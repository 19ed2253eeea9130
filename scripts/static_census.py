"""scripts/static_census.py — what the compiled kernels are, without a GPU: per kernel of
rejit_b200/librejit_b200.so the register / shared / local-memory use (cuobjdump -res-usage) and a census
of its SASS (cuobjdump -sass): TMA bulk copies (UBLKCP), mbarrier waits (SYNCS), global / shared loads,
shuffles, votes, local-memory traffic (spills or local arrays) and atomics.  Output: profiles/<tag>_static.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rejit_b200", "librejit_b200.so")


def demangle(name):
    n = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    n = re.sub(r"\(.*", "", n)
    n = n[5:] if n.startswith("void ") else n
    if "cub::" in n:
        m = re.search(r"(Device\w+Kernel|EmptyKernel)", n)
        n = "cub::" + (m.group(1) if m else n.split("::")[-1][:30])
    return n.replace("rejit_b200::", "")


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    rows = []
    for part in re.split(r"\n\s*Function : ", sass)[1:]:
        name = part.split("\n", 1)[0].strip()
        ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", part)
        head = collections.Counter(i.split(".")[0] for i in ins)
        full = collections.Counter(ins)
        tma = sum(v for k, v in full.items() if k.startswith(("UBLKCP", "UTMALDG", "UTMASTG")))
        atom = sum(head.get(k, 0) for k in ("ATOMS", "ATOMG", "ATOM", "RED"))
        reg, stack, shared, local = usage.get(name, (0, 0, 0, 0))
        rows.append((demangle(name), reg, stack, shared, len(ins), tma, head.get("SYNCS", 0), head.get("LDG", 0),
                     head.get("LDS", 0), head.get("STS", 0), head.get("SHFL", 0), head.get("VOTE", 0),
                     head.get("LDL", 0) + head.get("STL", 0), atom))
    fmt = "%-34s %4s %6s %7s %7s %4s %5s %5s %5s %5s %5s %5s %6s %5s"
    out = [fmt % ("kernel", "regs", "stack", "static$", "instrs", "TMA", "SYNCS", "LDG", "LDS", "STS", "SHFL", "VOTE",
                  "LDL+STL", "atom")]
    for r in sorted(rows, key=lambda r: -r[4]):
        out.append(fmt % ((r[0][:34],) + r[1:]))
    out.append("")
    out.append("static$ = static shared memory in bytes (the scan kernels use dynamic shared memory); TMA = UBLKCP / UTMALDG / "
               "UTMASTG; SYNCS = mbarrier operations; LDL+STL = local-memory instructions (NFA position sets and hit lists "
               "indexed at run time live there; `stack` is their frame).")
    text = "\n".join(out) + "\n"
    tag = sys.argv[1] if len(sys.argv) > 1 else "static"
    path = os.path.join(ROOT, "profiles", tag + "_static.txt")
    open(path, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
