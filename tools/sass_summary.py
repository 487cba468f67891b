"""SASS opcode summary of libzs.so (cuobjdump -sass), per kernel: the Blackwell tells (UTCHMMA = tcgen05.mma, LDTM =
tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = bulk copy, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops) plus the
instruction total.  Runs anywhere nvcc's cuobjdump is installed (no GPU needed):

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ossid_code_b200", "libzs.so")
TELLS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "UTCATOMSWS", "FFMA2", "FADD2", "FMNMX3")


def demangle_short(name):
    """_ZN...zs_k_mlp_tcILb1EEEv... -> zs_k_mlp_tc<true> (template arguments kept, namespace and parameters dropped)."""
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    m = re.search(r"(zs_k_[A-Za-z0-9_]+(?:<[^>]*>)?)", out)
    return m.group(1) if m else out


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(demangle_short(m.group(1)), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    arch = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout.strip().splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels; ELF images: {', '.join(a.split()[-1] for a in arch)}")
    print(f"# {'kernel':48s} {'instr':>6s} " + " ".join(f"{t:>8s}" for t in TELLS))
    total = collections.Counter()
    for name, c in kernels.items():
        total.update(c)
        print(f"  {name:48s} {sum(c.values()):6d} " + " ".join(f"{c.get(t, 0):8d}" for t in TELLS))
    print(f"  {'TOTAL':48s} {sum(total.values()):6d} " + " ".join(f"{total.get(t, 0):8d}" for t in TELLS))


if __name__ == "__main__":
    sys.exit(main())
