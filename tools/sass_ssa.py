#!/usr/bin/env python3
"""Turn a straight-line SASS listing (cuobjdump -sass, one kernel) into SSA form so that the
floating-point contraction tree chosen by ptxas can be read off reliably.

Usage: sass_ssa.py listing.txt [--expr vN ...] [--depth D]

Each register write gets a fresh value id vN; operands are printed as the id of their last
definition (64-bit operands use the id of the even register, tagged 'd').  Predicated writes
become  vN = sel(P, new, old).  Branches are ignored (fall-through), which is adequate for the
DXT kernels whose only branches guard slow paths (roundf / rcp fix-ups).

This is a *reading aid* for deriving the arithmetic contract of a compiled reference kernel; it
contains nothing of the reference itself.
"""
import re
import sys

WIDE = ("DADD", "DMUL", "DFMA", "F2F.F64.F32", "IMAD.WIDE", "DSETP")


def parse(path):
    out = []
    for line in open(path):
        m = re.match(r"\s*(?:\d+:\s*)?(?:/\*[0-9a-f]+\*/)?\s*(@!?U?P\d\s+)?([A-Z0-9_.]+)\s*(.*?)\s*;", line)
        if not m:
            continue
        pred, op, rest = m.group(1), m.group(2), m.group(3)
        ops = [o.strip() for o in re.split(r",\s*(?![^\[]*\])", rest)] if rest else []
        out.append((pred.strip() if pred else None, op, ops))
    return out


class SSA:
    def __init__(self):
        self.cur = {}      # reg -> vid
        self.defs = {}     # vid -> (op, [operand strings])
        self.n = 0

    def use(self, tok, wide=False):
        t = tok
        neg = ""
        m = re.match(r"^(-|\||-\||!)?(U?R\d+|U?P\d+|RZ|URZ|PT|UPT)(\.[A-Za-z0-9]+)*(\|)?(\.reuse)?$", t)
        if not m:
            return t  # immediate / const bank / address
        pre, reg = m.group(1) or "", m.group(2)
        suf = "".join(re.findall(r"\.(B\d|H\d(?:_H\d)?|X4|64)", t))
        if reg in ("RZ", "URZ", "PT", "UPT"):
            return pre + reg
        v = self.cur.get(reg)
        name = f"v{v}" if v is not None else f"in_{reg}"
        if wide:
            name += "d"
        if suf:
            name += "." + suf
        if "|" in t:
            name = "|" + name + "|"
            pre = pre.replace("|", "")
        return pre + name

    def define(self, reg, op, operands, pred=None):
        self.n += 1
        vid = self.n
        old = self.cur.get(reg)
        if pred:
            operands = [f"pred={pred}"] + operands + [f"old=v{old}" if old else "old=?"]
        self.defs[vid] = (op, operands, reg)
        self.cur[reg] = vid
        return vid


def main():
    args = sys.argv[1:]
    path = args[0]
    want = []
    depth = 6
    i = 1
    while i < len(args):
        if args[i] == "--expr":
            want.append(int(args[i + 1].lstrip("v")))
            i += 2
        elif args[i] == "--depth":
            depth = int(args[i + 1])
            i += 2
        else:
            i += 1
    ssa = SSA()
    lines = []
    for pred, op, ops in parse(path):
        if op in ("NOP", "BRA", "EXIT", "BSSY.RECONVERGENT", "BSYNC.RECONVERGENT", "RET.REL.NODEC", "CALL.REL.NOINC"):
            lines.append(f"      {pred or ''} {op} {' '.join(ops)}")
            continue
        wide = any(op.startswith(w) for w in WIDE)
        store = op.startswith("STG") or op.startswith("STS")
        if store or not ops:
            lines.append(f"      {pred or ''} {op} " + ", ".join(ssa.use(o) for o in ops))
            continue
        # destination(s): first operand; FSETP/ISETP/PLOP3 etc. write 2 predicates
        ndst = 2 if re.match(r"(FSETP|ISETP|DSETP|PLOP3|UISETP)", op) else 1
        if op.startswith("IADD3") or op.startswith("LEA") and len(ops) > 3 and ops[1].startswith("P"):
            pass
        dsts = ops[:ndst]
        srcs = ops[ndst:]
        # IADD3 / LEA may carry carry-out predicate operands right after dst
        extra = []
        while srcs and re.match(r"^!?U?P(T|\d)$", srcs[0]) and op.split(".")[0] in ("IADD3", "LEA", "UIADD3"):
            extra.append(srcs.pop(0))
        src_wide = wide and not op.startswith("F2F.F64.F32") and not op.startswith("IMAD.WIDE")
        if op.startswith("F2F.F32.F64"):
            src_wide = True
        use_ops = [ssa.use(s, wide=src_wide and re.match(r"^-?\|?R\d", s) is not None) for s in srcs]
        ppred = ssa.use(pred.lstrip("@")) if pred else None
        vids = []
        for d in dsts:
            if d in ("PT", "RZ", "UPT", "URZ"):
                continue
            vids.append(ssa.define(d.split(".")[0], op, use_ops, ppred))
        tag = ",".join(f"v{v}" for v in vids)
        lines.append(f"{tag:>10} = {op}({', '.join(use_ops)})" + (f"   if {ppred}" if ppred else "") + f"   [{','.join(dsts)}]")
    if not want:
        print("\n".join(lines))
        return

    def expr(v, d):
        tok = v
        m = re.match(r"^(-|\||-\|)?v(\d+)(d)?(\..*)?(\|)?$", tok)
        if not m or d == 0:
            return tok
        pre, vid = m.group(1) or "", int(m.group(2))
        op, ops, _ = ssa.defs[vid]
        inner = f"{op}({', '.join(expr(o, d - 1) for o in ops)})"
        return pre + inner + (m.group(4) or "")

    for w in want:
        print(f"v{w} = {expr('v%d' % w, depth)}")


if __name__ == "__main__":
    main()
