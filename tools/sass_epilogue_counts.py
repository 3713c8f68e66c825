"""Instruction counts of the hot kernel's epilogue variants, read off the SASS of the built object (no GPU needed).

The tcgen05 likelihood kernel is bound by the issue slots of its epilogue, so instructions per
(delay, bin) is the number to watch when the epilogue changes.  This walks `cuobjdump -sass` of
pyipn_b200/csrc/build/loglik_tc.o, cuts every `loglik_tc_kernel<7, MODE>` at its branch instructions and reports the
straight-line blocks that hold an 8-bin epilogue unit (recognised by the 16 int -> float conversions of the exponent assembly
and the 64+ FFMA of the polynomials), with the
opcode histogram that tells the variants apart (MUFU.LG2 = per-bin logarithm, DFMA = product epilogue, DADD/DMUL = fp64).

    python tools/sass_epilogue_counts.py [--md]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "pyipn_b200", "csrc", "build", "loglik_tc.o")
MODES = {0: "fp32 (per-bin logarithm)", 1: "compensated fp32", 2: "fp64", 3: "store (ipn_rff_eval)"}


def blocks(mode):
    name = f"_ZN3ipn16loglik_tc_kernelILi7ELi{mode}EEEvNS_8TcParamsE"
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", name, OBJ], check=True, capture_output=True, text=True).stdout
    ops = []
    for line in txt.splitlines():
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            ops.append(m.group(2))
    out, cur = [], []
    for op in ops + ["EXIT"]:
        if op.split(".")[0] in ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "CALL", "RET"):
            if cur:
                out.append(cur)
            cur = []
        else:
            cur.append(op)
    return len(ops), out


def describe(block):
    h = collections.Counter(op.split(".")[0] for op in block)
    if h.get("MUFU"):
        kind = "with counts: MUFU.LG2 + refinement per bin"
    elif h.get("DFMA"):
        kind = f"with counts: product, {h['DFMA'] // 8} multiplication(s) per bin"
    elif h.get("DADD") or h.get("DMUL"):
        kind = "fp64"
    else:
        kind = "no counts in the tile (only -kappa u)" if not h.get("ISETP") else "per-bin loop over the counts (prologue)"
    return kind, h


def main():
    md = "--md" in sys.argv
    if md:
        print("| Instantiation | Block | Instructions per 8 bins | per (delay, bin) | FFMA | FADD | FMUL | MUFU | F2F | DFMA | integer / other |")
        print("|---|---|---|---|---|---|---|---|---|---|---|")
    for mode, label in MODES.items():
        try:
            total, bl = blocks(mode)
        except subprocess.CalledProcessError:
            continue
        seen = set()
        for b in bl:
            h = collections.Counter(op.split(".")[0] for op in b)
            # an 8-bin unit: eight I2FP pairs of the integer exponent assembly (16 conversions) and eight polynomials (>= 64 FFMA)
            if h.get("I2FP", 0) < 16 or h.get("FFMA", 0) < 64 or len(b) > 700:
                continue
            kind, h = describe(b)
            key = (kind, len(b))
            if key in seen:
                continue
            seen.add(key)
            other = len(b) - sum(h.get(k, 0) for k in ("FFMA", "FADD", "FMUL", "MUFU", "F2F", "DFMA"))
            if md:
                print(f"| `<7, {mode}>` {label} ({total} instructions) | {kind} | {len(b)} | {len(b) / 8:.1f} | {h.get('FFMA', 0)} | "
                      f"{h.get('FADD', 0)} | {h.get('FMUL', 0)} | {h.get('MUFU', 0)} | {h.get('F2F', 0)} | {h.get('DFMA', 0)} | {other} |")
            else:
                print(f"<7,{mode}> {label:28s} {kind:52s} {len(b):4d} per 8 bins = {len(b) / 8:5.1f} per bin   {dict(h)}")


if __name__ == "__main__":
    main()
