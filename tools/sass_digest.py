#!/usr/bin/env python
"""Per-kernel digest of the SASS inside libmctq_sm100.so (runs anywhere cuobjdump is installed, no GPU needed):

    python tools/sass_digest.py > profiles/rNN_sass_digest.md

For every kernel family: number of instantiations, registers (min-max), and how often the mnemonics that carry the design
appear -- 128 / 256-bit streaming loads and stores (LDG.E.128 / STG.E.128 / STG.E.ENL2.256), bulk-copy staging (UBLKCP, SYNCS =
mbarrier), shared-memory look-ups (LDS), dependent launch (ACQBULK / PREEXIT are what griddepcontrol.wait /
launch_dependents compile to), packed half-precision math, and the absence of tensor-core instructions (HMMA / UTCMMA).
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mct_quantizers_b200", "libmctq_sm100.so")
PATTERNS = [("LDG.128", r"\bLDG\.E(\.[A-Z0-9_]+)*\.128"), ("LDG.64", r"\bLDG\.E(\.[A-Z0-9_]+)*\.64"), ("STG.128", r"\bSTG\.E(\.[A-Z0-9_]+)*\.128"),
            ("STG.256", r"\bSTG\.E(\.[A-Z0-9_]+)*\.256"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"),
            ("SHFL", r"\bSHFL"), ("FFMA", r"\bFFMA"), ("FMUL", r"\bFMUL"), ("FADD", r"\bFADD"), ("HFMA2/HADD2", r"\bH(FMA|ADD)2"),
            ("MUFU.RCP", r"\bMUFU\.RCP"), ("ACQBULK", r"\bACQBULK"), ("PREEXIT", r"\bPREEXIT"), ("LDC", r"\bU?LDC"), ("BAR", r"\bBAR\."),
            ("HMMA/UTCMMA", r"\b(HMMA|UTCHMMA|UTCQMMA|UTCMMA)")]


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+)", ln)
        if m and cur:
            regs[cur] = int(m.group(1))
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fam = collections.OrderedDict()
    cur = None
    arch = set(re.findall(r"arch = (sm_\w+)", sass))
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            dem = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"^void ", "", dem)
            name = re.sub(r"<.*", "", name).replace("mctq::", "").split("(")[0]
            f = fam.setdefault(name, {"n": 0, "regs": [], "counts": collections.Counter(), "instr": 0})
            f["n"] += 1
            if cur in regs:
                f["regs"].append(regs[cur])
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            f = fam[name]
            f["instr"] += 1
            for key, pat in PATTERNS:
                if re.search(pat, ln):
                    f["counts"][key] += 1
    print(f"# SASS digest of libmctq_sm100.so ({', '.join(sorted(arch)) or 'sm_100a'}; cuobjdump -sass, summed over all instantiations of a family)\n")
    keys = [k for k, _ in PATTERNS]
    print("| kernel family | inst. | regs | SASS instr | " + " | ".join(keys) + " |")
    print("|---|---:|---:|---:|" + "---:|" * len(keys))
    for name, f in fam.items():
        r = f"{min(f['regs'])}-{max(f['regs'])}" if f["regs"] else "?"
        print(f"| `{name}` | {f['n']} | {r} | {f['instr']} | " + " | ".join(str(f["counts"].get(k, 0)) for k in keys) + " |")
    tot = collections.Counter()
    for f in fam.values():
        tot.update(f["counts"])
    print(f"\nTensor-core instructions in the whole library: {tot.get('HMMA/UTCMMA', 0)} (the path is HBM-bound elementwise work).")
    print(f"Bulk-copy (1-D TMA) instructions: {tot.get('UBLKCP', 0)}; 256-bit stores: {tot.get('STG.256', 0)}; 128-bit loads: {tot.get('LDG.128', 0)}.")


if __name__ == "__main__":
    main()
