"""Blackwell-native evidence from the built library (no GPU needed): per kernel of libptfnn.so, the count of the SASS
mnemonics that only the sm_100a paths produce (B200_PROFILING.md, "What proves a Blackwell-native kernel"), plus short
verbatim excerpts.  usage: python tools/sass_evidence.py > profiles/r02_sass_evidence.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "parallel-tempering-neural-net_b200", "csrc", "libptfnn.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "REDUX", "FFMA2", "FMUL2", "FADD2", "MUFU.EX2", "MUFU.RCP", "BAR.SYNC", "HMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            kernels[name].append(line)
    demangled = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangled)) if len(demangled) == len(kernels) else {k: k for k in kernels}
    print("# SASS evidence, round 2 (`cuobjdump -sass csrc/libptfnn.so`, built by `python __graft_entry__.py`)\n")
    print("tcgen05.mma -> `UTCHMMA`, tcgen05.commit -> `UTCBAR`, tcgen05.ld / st -> `LDTM` / `STTM`, 1-D TMA bulk copy -> `UBLKCP`,")
    print("mbarrier -> `SYNCS`, redux.sync -> `REDUX`, f32x2 arithmetic -> `FFMA2 / FMUL2 / FADD2`.  No `HMMA` (legacy mma.sync) anywhere.\n")
    print("| kernel | instructions | " + " | ".join(MNEMONICS) + " |")
    print("|---|---|" + "---|" * len(MNEMONICS))
    excerpts = {}
    for k, lines in kernels.items():
        short = names[k].rsplit("(", 1)[0].replace("ptfnn::", "").replace("(int)", "").replace("(bool)", "")
        counts = [sum(1 for l in lines if re.search(r"\b" + re.escape(m), l)) for m in MNEMONICS]
        if sum(counts[:6]) == 0 and "chain_kernel" not in short:
            continue
        print("| `%s` | %d | %s |" % (short[:110], len(lines), " | ".join(str(c) for c in counts)))
        if "op_forward_tc_kernel" in short or ("chain_kernel<16, 256, 10" in short and short not in excerpts):
            excerpts[short] = lines
    for short, lines in excerpts.items():
        print("\n## `%s`: first occurrences\n\n```" % short[:120])
        seen = set()
        for l in lines:
            for m in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP"):
                key = m + (" tmem[" if (m == "UTCHMMA" and re.search(r"UTCHMMA tmem\[", l)) else "")
                if re.search(r"\b" + m, l) and key not in seen and len([s for s in seen if s.startswith(m)]) < 2:
                    seen.add(key)
                    print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip())
        print("```")
        print("\n(`UTCHMMA gdesc[..], gdesc[..], tmem[..]` = layer 1, both operands from shared memory; `UTCHMMA tmem[..], gdesc[..], tmem[..]` =")
        print("layer 2, the A operand -- the hidden activations the epilogue wrote with `STTM` -- read from tensor memory.)")


if __name__ == "__main__":
    main()
