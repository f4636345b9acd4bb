#!/usr/bin/env python3
"""Offline (no GPU) statistics of a kernel's SASS: instruction count, opcode histogram, sum of the static stall counts of the control codes,
and the sequence of pipe classes (F = FMA pipe, A = ALU pipe, D = FP64, X = conversion/special-function pipe, M = memory, . = other) - the
view that showed the long FMA-only / ALU-only stretches of the DXT1 encode and the one used to count instructions while
rewriting the DXT5-YCoCg index selection.

    python tools/sass_stats.py ultragrid_b200/csrc/dxt_kernels.o dxt_uyvy_kernelILi6ELi1ELb0      (substring of the mangled name)
"""
import collections
import re
import subprocess
import sys

PIPE = [
    (r"^(FFMA2|FFMA|FMUL2|FMUL|FADD2|FADD|HFMA2|HADD2|HMUL2|IMAD|IMUL|FSWZADD)", "F"),
    (r"^(FMNMX3|FMNMX|PRMT|LEA|LOP3|PLOP3|IADD3|VIADD|VIADDMNMX|VIMNMX3|VIMNMX|FSEL|SEL|FSETP|ISETP|FSET|SHF|SHL|SHR|MOV|IABS|LOP|POPC|BREV|FLO|BMSK|SGXT|I2FP|FCHK)", "A"),
    (r"^(DADD|DFMA|DMUL|DSETP|DMNMX)", "D"),
    (r"^(F2F|F2I|I2F|MUFU|FRND)", "X"),
    (r"^(LDG|STG|LDS|STS|LDL|STL|LDC|LDCU|ATOM|ATOMS|ATOMG|RED|LDGSTS|LDSM|S2R|S2UR|CS2R|UBLKCP)", "M"),
]


def pipe_class(op):
    for pat, c in PIPE:
        if re.match(pat, op):
            return c
    return "."


def main():
    obj, pat = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.split("Function : ")
    for f in txt[1:]:
        lines = f.split("\n")
        name = lines[0].strip()
        if pat not in name:
            continue
        ops, seq, stall, n = collections.Counter(), [], 0, 0
        for i, line in enumerate(lines):
            m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+).*?/\* (0x[0-9a-f]{16}) \*/", line)
            if not m:
                continue
            op = m.group(1).rstrip(";")
            ops[op.split(".")[0] if not op.startswith("F2F") else op] += 1
            seq.append(pipe_class(op))
            m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", lines[i + 1]) if i + 1 < len(lines) else None
            if m2:  # control bits live in the upper word of the 128-bit instruction: stall count = bits 105..108
                stall += (int(m2.group(1), 16) >> 41) & 0xF
            n += 1
        print(f"{name}\n  {n} instructions, static stall sum {stall}")
        print("  " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(24)))
        by_pipe = collections.Counter(seq)
        print("  by pipe class: " + "  ".join(f"{k}:{v}" for k, v in by_pipe.most_common()))
        s = "".join(seq)
        for i in range(0, len(s), 120):
            print(f"  {i:5d} {s[i:i + 120]}")


if __name__ == "__main__":
    try:
        main()
    except BrokenPipeError:  # output piped into head
        pass
