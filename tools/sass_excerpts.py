"""Regenerates profiles/r02_sass_excerpts.txt: per-kernel static SASS instruction mix of libflyby_b200.so
(cuobjdump -sass) with the mnemonics that back the design claims (UBLKCP = cp.async.bulk staging of the LBVH top,
SYNCS = mbarrier, ATOMG / RED = integer atomics, MATCH = warp match votes, DFMA/DMUL/DADD = the exact chain).
    python tools/sass_excerpts.py > profiles/r02_sass_excerpts.txt"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "fly-by-cnn_b200", "flyby_b200", "lib", "libflyby_b200.so")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout.split("\n")
    print("cuobjdump -sass fly-by-cnn_b200/flyby_b200/lib/libflyby_b200.so (sm_100a): static instruction mix per kernel.")
    print("  UBLKCP            bulk async copy global -> shared (cp.async.bulk, the TMA engine) staging the top of the LBVH")
    print("  SYNCS.*           mbarrier init / arrive-expect-tx / try-wait around it")
    print("  ATOMG.*.MIN.64    the 64-bit front atomicMin of the scan conversion; ATOMG / RED .ADD: integer vote atomics")
    print("  MATCH.ANY         warp match votes (radix-sort ranking, vote aggregation)")
    print("  DFMA/DMUL/DADD    the double-precision exact chain;  MUFU.RCP64H: seeds of the IEEE divisions / square root")
    print("  LDL / STL         local-memory (spill) accesses")
    print()
    fn, counts, order = None, {}, []
    pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)")
    for line in sass:
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            counts[fn] = {}
            order.append(fn)
            continue
        m = pat.match(line)
        if fn and m:
            counts[fn][m.group(1)] = counts[fn].get(m.group(1), 0) + 1
    names = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.split("\n")
    keys = ["UBLKCP", "SYNCS", "ATOMG", "RED", "MATCH", "DFMA", "DMUL", "DADD", "MUFU.RCP64H", "FFMA", "LDL", "STL"]
    for fn, name in zip(order, names):
        c = counts[fn]
        row = {k: sum(v for op, v in c.items() if op.startswith(k)) for k in keys}
        print("%-100s insts=%-6d %s" % (name[:100], sum(c.values()), " ".join("%s=%d" % kv for kv in row.items() if kv[1])))
        ex = [op for op in c if op.startswith(("UBLKCP", "ATOMG", "RED.", "MATCH", "SYNCS"))]
        if ex:
            print("       " + ", ".join("%s x%d" % (op, c[op]) for op in sorted(ex)))


if __name__ == "__main__":
    main()
