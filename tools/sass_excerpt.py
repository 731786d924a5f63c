"""Developer tool (no GPU needed): memory / synchronisation instructions of a kernel's SASS for profiles/.

    python tools/sass_excerpt.py <mangled kernel name> <title> <out file>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
KEY = re.compile(r"UBLKCP|SYNCS|FENCE|LDG|ATOMG|STG|LDS|SHFL|DFMA|WARPSYNC|NANOSLEEP|VOTE|ELECT|REDUX|CCTL|UTMA")


def main():
    fun, title, out = sys.argv[1:4]
    obj = os.path.join(ROOT, "hifir_b200", "_lib", "obj", "wsweep.o")
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True, check=True).stdout
    ins = [l.rstrip() for l in sass.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    hist, sel = collections.Counter(), []
    for l in ins:
        m = re.search(r"\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m and KEY.search(m.group(2)):
            hist[m.group(2)] += 1
            sel.append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l))
    head = [f"# SASS excerpt: {title}",
            "# cuobjdump -sass -fun <mangled> hifir_b200/_lib/obj/wsweep.o (sm_100a, nvcc 12.9, -O3 -lineinfo); memory / synchronisation",
            "# instructions only.  TMA bulk copy = UBLKCP.S.G; mbarrier = SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK.TRANS64.TRYWAIT;",
            "# polling gathers = LDG.E.64/128.STRONG.GPU; publish = ATOMG.E.EXCH.64.STRONG.GPU / STG.E.128.STRONG.GPU; factor entries",
            "# from the ring stage = LDS / LDS.64",
            f"# {len(ins)} instructions in the kernel; histogram of the selected mnemonics:",
            "# " + ", ".join(f"{k} x{v}" for k, v in hist.most_common())]
    open(out, "w").write("\n".join(head + sel) + "\n")
    print(out, len(ins), "instructions,", len(sel), "selected")


if __name__ == "__main__":
    main()
