"""SASS evidence for the two iteration kernels (no GPU needed): instruction mix of the tile loops,
TMA / mbarrier / cluster / DSMEM mnemonics, and where local-memory (spill) accesses sit.
python scripts/sass_report.py > profiles/sass_gn_iteration_ndt6.txt"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "nonlinear_optimizer_for_slam_b200", "csrc", "nlo_kernels.o")
KERNELS = [
    ("streaming kernel, ndt6 / Exponential / fp64 storage, two warp groups (default)  (gn_iteration_kernel<0,1,double,2>)",
     "_ZN3nlo19gn_iteration_kernelILi0ELi1EdLi2EEEvNS_10IterParamsE"),
    ("streaming kernel, ndt6 / Exponential / fp64 storage, one warp group  (gn_iteration_kernel<0,1,double,1>)",
     "_ZN3nlo19gn_iteration_kernelILi0ELi1EdLi1EEEvNS_10IterParamsE"),
    ("resident kernel, ndt6 / Exponential  (gn_resident_kernel<0,1>)",
     "_ZN3nlo18gn_resident_kernelILi0ELi1EEEvNS_10IterParamsE"),
]


def sass(fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, OBJ], capture_output=True, text=True, check=True).stdout
    return [l for l in out.splitlines() if re.match(r"^\s+/\*[0-9a-f]{4,5}\*/", l)]


def mnemonic(line):
    m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    return m.group(1) if m else ""


def main():
    for title, fun in KERNELS:
        lines = sass(fun)
        ops = [mnemonic(l) for l in lines]
        print("=" * 100)
        print(title)
        print("instructions: %d (%.1f KB)" % (len(lines), len(lines) * 16 / 1024.0))
        c = Counter(o.split(".")[0] for o in ops)
        keys = ["UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "LDS", "SHFL", "MUFU", "LDL", "STL", "BAR", "UCGABAR_ARV",
                "UCGABAR_WAIT", "MAPA", "CALL", "ATOM", "ATOMG", "RED", "MEMBAR", "CCTL"]
        print("whole kernel: " + "  ".join("%s %d" % (k, c.get(k, 0)) for k in keys))
        print("cluster barrier / LL words (relaxed gpu- and sys-scope accesses; the DSMEM stores are the ST.E...STRONG.GPU): "
              + "  ".join("%s %d" % (k, sum(1 for o in ops if o.startswith(k)))
                          for k in ["UCGABAR", "ST.E.128.STRONG.GPU", "STG.E.128.STRONG.GPU", "LDG.E.128.STRONG.GPU",
                                    "STG.E.128.STRONG.SYS", "LDG.E.128.STRONG.SYS", "LDG.E.64.STRONG.GPU", "LDG.E.STRONG.GPU"]))
        # tile loop = from the last full-barrier wait (or first LDS.64 block) before the first SHFL to that SHFL
        first_shfl = next(i for i, o in enumerate(ops) if o.startswith("SHFL"))
        waits = [i for i, o in enumerate(ops) if o.startswith("SYNCS.PHASECHK") and i < first_shfl]
        dfma = [i for i, o in enumerate(ops) if o.startswith("DFMA") and i < first_shfl]
        start = waits[-1] if waits else dfma[0] - 40
        body = ops[start:first_shfl]
        cb = Counter(o.split(".")[0] for o in body)
        print("tile loop (instructions %d..%d): %d instructions: %s" %
              (start, first_shfl, len(body), "  ".join("%s %d" % kv for kv in cb.most_common(14))))
        print("  local-memory accesses in the tile loop: LDL %d, STL %d" % (cb.get("LDL", 0), cb.get("STL", 0)))
        spill = [i for i, o in enumerate(ops) if o.startswith("LDL") or o.startswith("STL")]
        print("  local-memory accesses elsewhere (instruction indices): %s" %
              (", ".join(str(i) for i in spill if not start <= i < first_shfl)[:400] or "none"))
        print("tile loop excerpt (first 60 instructions):")
        for l in lines[start:start + 60]:
            print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l.rstrip()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
