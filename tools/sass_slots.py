#!/usr/bin/env python3
"""Static pipe-slot model of a SASS region (tuning aid).

usage: sass_slots.py <sass file> [<start hex> <end hex> [<repeat>]]...
Slot costs follow the B200 measurements in profiles/pipes_r1.jsonl: zero-addend IMAD.WIDE / IMAD / IADD3 / LOP3 /
SHF / SEL issue at 64 lanes/clk/SM (1 slot); IMAD.WIDE with a register addend and IMAD.HI at 32 (2 slots).
"""
import re
import sys

LINE = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)\s*(.*?);")


def classify(op, args):
    base = op.split(".")[0]
    if base == "IMAD":
        if ".WIDE" in op:
            addend = args.split(",")[-1].strip()
            if ".X" in op:
                addend = args.split(",")[-2].strip()
            return ("fma", 1 if addend == "RZ" else 2)
        if ".HI" in op:
            return ("fma", 2)
        return ("fma", 1)
    if base in ("HFMA2", "FFMA", "FMUL", "FADD", "VIADD"):
        return ("fma", 1)
    if base in ("IADD3", "LOP3", "SHF", "SEL", "ISETP", "MOV", "PRMT", "LEA", "IABS", "IMNMX", "PLOP3", "P2R", "R2P", "CS2R", "VIMNMX"):
        return ("alu", 1)
    if base.startswith("U") or base in ("R2UR",):
        return ("uni", 1)
    return ("other", 1)


def region(lines, lo, hi):
    tot = {"fma": 0, "alu": 0, "uni": 0, "other": 0, "instr": 0}
    for addr, op, args in lines:
        if lo <= addr < hi:
            k, c = classify(op, args)
            tot[k] += c
            tot["instr"] += 1
    return tot


def main():
    lines = []
    for ln in open(sys.argv[1]):
        m = LINE.match(ln)
        if m:
            lines.append((int(m.group(1), 16), m.group(2), m.group(3)))
    spec = sys.argv[2:]
    if not spec:
        spec = ["0", "fffff", "1"]
    grand = {"fma": 0, "alu": 0, "uni": 0, "other": 0, "instr": 0}
    i = 0
    while i < len(spec):
        lo, hi = int(spec[i], 16), int(spec[i + 1], 16)
        rep = int(spec[i + 2]) if i + 2 < len(spec) and not spec[i + 2].startswith("0x") else 1
        t = region(lines, lo, hi)
        print("%05x-%05x x%-3d instr=%5d fma=%5d alu=%5d uni=%4d other=%4d" % (lo, hi, rep, t["instr"], t["fma"], t["alu"], t["uni"], t["other"]))
        for k in grand:
            grand[k] += rep * t[k]
        i += 3
    cyc = max(2 * grand["fma"], 2 * grand["alu"], grand["instr"])
    print("TOTAL instr=%d fma_slots=%d alu_slots=%d uni=%d other=%d -> >= %d cycles/warp/SMSP; ideal %.3f Gperm/s @1.92GHz" % (
        grand["instr"], grand["fma"], grand["alu"], grand["uni"], grand["other"], cyc, 148 * 4 * 32 * 1.92e9 / cyc / 1e9))


if __name__ == "__main__":
    main()
