#!/usr/bin/env python3
"""Generate lightning_b200/csrc/u256_gen.cuh: the inline-PTX bodies of the 256x256->512 product and the
256-bit square for sm_100a.

Layout (same idea as the hand-written first version, see u256.cuh): two accumulator banks, E for
64-bit products whose low limb sits on an even limb position, O for odd positions (O[k] is limb k+1).
Every `mad.lo.cc/madc.hi.cc` pair on one (a,b) becomes ONE IMAD.WIDE.U32(.X) with the carry in a
predicate.  The generator tracks which limbs already hold data so that
  * the first product landing on a limb is a plain multiply (no zero-initialised registers),
  * a chain that ends on untouched limbs needs no carry-out,
  * a chain that ends on live limbs captures its carry straight into the (fresh) next limb.
Run:  python tools/gen_mul.py > lightning_b200/csrc/u256_gen.cuh
"""
import sys


class Asm:
    """one asm() statement: PTX lines with named operands, resolved to %n at emit time"""

    def __init__(self):
        self.lines, self.outs, self.ins, self.inouts = [], [], [], []

    def ref(self, name, mode):
        lst = {"out": self.outs, "in": self.ins, "io": self.inouts}[mode]
        # an operand that is written must not also be listed as plain input
        if mode == "in" and (name in self.outs or name in self.inouts):
            return "{" + name + "}"
        if mode in ("out", "io") and name in self.ins:
            self.ins.remove(name)
        if mode == "io" and name in self.outs:
            return "{" + name + "}"  # written earlier in this block; still the same register
        if name not in lst:
            lst.append(name)
        return "{" + name + "}"

    def emit(self, indent="    "):
        order = self.inouts + self.outs + self.ins
        idx = {n: i for i, n in enumerate(order)}
        body = []
        for l in self.lines:
            for n in order:
                l = l.replace("{" + n + "}", "%" + str(idx[n]))
            body.append(l)
        s = indent + 'asm("' + ('\\n\\t"\n' + indent + '    "').join(body) + '"\n'
        cons = [f'"+r"({n})' for n in self.inouts] + [f'"=r"({n})' for n in self.outs]
        s += indent + "    : " + ", ".join(cons) + "\n"
        s += indent + "    : " + ", ".join(f'"r"({n})' for n in self.ins) + ");\n"
        assert len(order) <= 60
        return s


def gen_chain(bank, live, start, prods):
    """Emit one chain.  If the top limb written was already live, capture the carry into the next limb."""
    top = start + 2 * len(prods) - 1
    top_was_live = top in live
    lo_first_fresh = start not in live
    a = Asm()
    started = False  # carry chain started
    n = len(prods)
    for k, (x, y) in enumerate(prods):
        for part, limb in (("lo", start + 2 * k), ("hi", start + 2 * k + 1)):
            name = f"{bank}[{limb}]"
            is_live = limb in live
            last = (k == n - 1 and part == "hi")
            if not started and not is_live:
                a.lines.append(f"mul.{part}.u32 {a.ref(name, 'out')}, {a.ref(x, 'in')}, {a.ref(y, 'in')};")
            else:
                cin = "c" if started else ""
                cout = ".cc" if (not last or top_was_live) else ""
                if is_live:
                    r = a.ref(name, "io")
                    a.lines.append(f"mad{cin}.{part}{cout}.u32 {r}, {a.ref(x, 'in')}, {a.ref(y, 'in')}, {r};")
                else:
                    a.lines.append(f"mad{cin}.{part}{cout}.u32 {a.ref(name, 'out')}, {a.ref(x, 'in')}, {a.ref(y, 'in')}, 0;")
                started = True
    for k in range(n):
        live.add(start + 2 * k)
        live.add(start + 2 * k + 1)
    if top_was_live:
        nxt = top + 1
        assert nxt not in live, (bank, nxt)
        a.lines.append(f"addc.u32 {a.ref(f'{bank}[{nxt}]', 'out')}, 0, 0;")
        live.add(nxt)
    return a.emit()


def gen_merge(n_limbs, e_live, o_live, dst="r"):
    """dst[0] = E[0]; dst[k] = E[k] + O[k-1] + carry"""
    a = Asm()
    out = f"    {dst}[0] = E[0];\n"
    first = True
    for k in range(1, n_limbs):
        e = a.ref(f"E[{k}]", "in") if k in e_live else "0"
        o = a.ref(f"O[{k-1}]", "in") if (k - 1) in o_live else "0"
        last = k == n_limbs - 1
        op = ("add" if first else "addc") + ("" if last else ".cc") + ".u32"
        a.lines.append(f"{op} {a.ref(f'{dst}[{k}]', 'out')}, {e}, {o};")
        first = False
    return out + a.emit()


def gen_mul():
    out = []
    out.append("// r[0..15] = a[0..7] * b[0..7]   (64 IMAD.WIDE.U32)")
    out.append("SV_D void sv_mul8_dev(u32* __restrict__ r, const u32* __restrict__ a, const u32* __restrict__ b) {")
    out.append("    u32 E[16], O[16];")
    e_live, o_live = set(), set()
    for i in range(8):
        ev = [(f"a[{j}]", f"b[{i}]") for j in (0, 2, 4, 6)]  # a_even * b_i -> position i + j (parity of i)
        od = [(f"a[{j}]", f"b[{i}]") for j in (1, 3, 5, 7)]  # a_odd  * b_i -> position i + j (parity of i+1)
        if i % 2 == 0:
            out.append(gen_chain("E", e_live, i, ev).rstrip("\n"))
            out.append(gen_chain("O", o_live, i, od).rstrip("\n"))  # position i+1 -> O index i
        else:
            out.append(gen_chain("O", o_live, i - 1, ev).rstrip("\n"))  # position i -> O index i-1
            out.append(gen_chain("E", e_live, i + 1, od).rstrip("\n"))
    out.append(gen_merge(16, e_live, o_live).rstrip("\n"))
    out.append("}")
    return "\n".join(out)


def gen_sqr():
    out = []
    out.append("// r[0..15] = a[0..7]^2   (28 doubled cross products + 8 squares = 36 IMAD.WIDE.U32)")
    out.append("SV_D void sv_sqr8_dev(u32* __restrict__ r, const u32* __restrict__ a) {")
    out.append("    u32 E[16], O[16], S[16], D[16], T[16];")
    e_live, o_live = set(), set()
    for i in range(7):
        js = list(range(i + 1, 8))
        # position i + j ; odd positions -> O[pos-1], even -> E[pos]
        odd_pos = [j for j in js if (i + j) % 2 == 1]
        even_pos = [j for j in js if (i + j) % 2 == 0]
        if odd_pos:
            out.append(gen_chain("O", o_live, i + odd_pos[0] - 1, [(f"a[{i}]", f"a[{j}]") for j in odd_pos]).rstrip("\n"))
        if even_pos:
            out.append(gen_chain("E", e_live, i + even_pos[0], [(f"a[{i}]", f"a[{j}]") for j in even_pos]).rstrip("\n"))
    # cross sum S (limb 0 and limb 1's E part are empty)
    a = Asm()
    first = True
    lines = ["    S[0] = 0;"]
    for k in range(1, 16):
        e = a.ref(f"E[{k}]", "in") if k in e_live else "0"
        o = a.ref(f"O[{k-1}]", "in") if (k - 1) in o_live else "0"
        last = k == 15
        op = ("add" if first else "addc") + ("" if last else ".cc") + ".u32"
        a.lines.append(f"{op} {a.ref(f'S[{k}]', 'out')}, {e}, {o};")
        first = False
    out.append("\n".join(lines))
    out.append(a.emit().rstrip("\n"))
    # diagonal
    for half in (0, 1):
        a = Asm()
        for i in range(4 * half, 4 * half + 4):
            a.lines.append(f"mul.lo.u32 {a.ref(f'D[{2*i}]', 'out')}, {a.ref(f'a[{i}]', 'in')}, {a.ref(f'a[{i}]', 'in')};")
            a.lines.append(f"mul.hi.u32 {a.ref(f'D[{2*i+1}]', 'out')}, {a.ref(f'a[{i}]', 'in')}, {a.ref(f'a[{i}]', 'in')};")
        out.append(a.emit().rstrip("\n"))
    # r = D + S + S  (S[0] == 0)
    for dst, x in (("T", "D"), ("r", "T")):
        a = Asm()
        out.append(f"    {dst}[0] = {x}[0];")
        for k in range(1, 16):
            op = ("add" if k == 1 else "addc") + ("" if k == 15 else ".cc") + ".u32"
            a.lines.append(f"{op} {a.ref(f'{dst}[{k}]', 'out')}, {a.ref(f'{x}[{k}]', 'in')}, {a.ref(f'S[{k}]', 'in')};")
        out.append(a.emit().rstrip("\n"))
    out.append("}")
    return "\n".join(out)


if __name__ == "__main__":
    print("// u256_gen.cuh — GENERATED by tools/gen_mul.py; do not edit.  Device-only PTX bodies.")
    print("#pragma once")
    print('#include "common.cuh"')
    print("#if SV_DEVICE_CODE")
    print(gen_mul())
    print()
    print(gen_sqr())
    print("#endif")
