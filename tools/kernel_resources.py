"""Static evidence about the built kernels, produced without a GPU: registers / stack (spills) / shared memory per kernel
(`cuobjdump -res-usage`) and the SASS memory-instruction mix of the kernels the roofline accounting is about
(`cuobjdump -sass`).  Usage: python tools/kernel_resources.py > profiles/kernel_resources_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ufemism_b200", "libufemism_b200.so")
# kernels whose instruction mix is listed (substring of the demangled name)
SASS_OF = ["k_ssa_sor<true, true, false>", "k_ssa_viscosity<false, 4>", "k_geom_ac", "k_thk<1>", "k_sia_ac("]
MEM = re.compile(r"\b(LDG|STG|LDS|STS|LDL|STL|RED|ATOMG|ATOMS|UBLKCP|SYNCS|LDGSTS|SHFL|BAR|MEMBAR|CCTL|ERRBAR|LDGDEPBAR|DEPBAR|MUFU|DFMA|DADD|DMUL)(\.[A-Z0-9_.]+)?")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.split("\n")
    return dict(zip(names, out))


def resources():
    txt = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    rows, fn = [], None
    for line in txt.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        if fn and "REG:" in line:
            d = dict(re.findall(r"([A-Z]+(?:\[\d\])?):(\d+)", line))
            rows.append((fn, int(d["REG"]), int(d.get("STACK", 0)), int(d.get("SHARED", 0)), int(d.get("LOCAL", 0))))
            fn = None
    return rows


def sass_mix():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    mix, fn = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            mix[fn] = collections.Counter()
            continue
        if fn:
            m = MEM.search(line)
            if m and "/*" in line:
                mix[fn][m.group(0)] += 1
    return mix


def main():
    arch = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True, check=True).stdout.strip()
    print("# cubins in libufemism_b200.so (sm_100a only)\n" + arch + "\n")
    rows = resources()
    names = demangle([r[0] for r in rows])
    print("# registers / stack bytes (spills + local arrays) / static shared bytes / local, per kernel (cuobjdump -res-usage)")
    print(f"{'REG':>4} {'STACK':>6} {'SHARED':>7} {'LOCAL':>6}  kernel")
    for fn, reg, stack, sh, loc in sorted(rows, key=lambda r: names[r[0]]):
        print(f"{reg:4d} {stack:6d} {sh:7d} {loc:6d}  {re.sub(r'\(.*', '', names[fn])}")
    mix = sass_mix()
    names = demangle(list(mix))
    print("\n# SASS instruction mix (static counts, memory / synchronisation / fp64 mnemonics) of the kernels of DESIGN.md section 4")
    for fn, c in mix.items():
        if any(k in names[fn] for k in SASS_OF):
            print(f"\n{re.sub(r'\(.*', '', names[fn])}")
            for k, v in sorted(c.items()):
                print(f"  {v:5d}  {k}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
