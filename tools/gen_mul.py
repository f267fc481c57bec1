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


# ---------------------------------------------------------------------------------------------------------------
# Programs: a list of ("asm", Asm) / ("c", "dst = src;") items that can be emitted as CUDA or SIMULATED in Python
# (PTX carry-flag semantics), so that new generated bodies are checked against big-int arithmetic before they
# ever reach a GPU (python tools/gen_mul.py --check).
# ---------------------------------------------------------------------------------------------------------------
def chain_asm(bank, live, start, prods):
    """like gen_chain, returning the Asm object"""
    top = start + 2 * len(prods) - 1
    top_was_live = top in live
    a = Asm()
    started = False
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
    return a


def prog_rect(prog, E, O, dst, a_names, b_names):
    """dst[0 .. na+nb) = a * b with the even/odd-bank layout (banks E, O are fresh array names)"""
    na, nb = len(a_names), len(b_names)
    e_live, o_live = set(), set()
    for i in range(nb):
        ev = [(a_names[j], b_names[i]) for j in range(0, na, 2)]
        od = [(a_names[j], b_names[i]) for j in range(1, na, 2)]
        if i % 2 == 0:
            prog.append(("asm", chain_asm(E, e_live, i, ev)))
            if od:
                prog.append(("asm", chain_asm(O, o_live, i, od)))
        else:
            prog.append(("asm", chain_asm(O, o_live, i - 1, ev)))
            if od:
                prog.append(("asm", chain_asm(E, e_live, i + 1, od)))
    n = na + nb
    prog.append(("c", f"{dst}[0] = {E}[0];"))
    a = Asm()
    first = True
    for k in range(1, n):
        e = a.ref(f"{E}[{k}]", "in") if k in e_live else "0"
        o = a.ref(f"{O}[{k-1}]", "in") if (k - 1) in o_live else "0"
        last = k == n - 1
        op = ("add" if first else "addc") + ("" if last else ".cc") + ".u32"
        a.lines.append(f"{op} {a.ref(f'{dst}[{k}]', 'out')}, {e}, {o};")
        first = False
    prog.append(("asm", a))


def prog_add(prog, dst, x, y, carry_out=None):
    """dst[i] = x[i] + y[i] (lists of operand names or "0", equal length) in one carry chain; optional carry-out name"""
    a = Asm()
    n = len(dst)
    for i in range(n):
        last = (i == n - 1) and carry_out is None
        op = ("add" if i == 0 else "addc") + ("" if last else ".cc") + ".u32"
        xs = a.ref(x[i], "in") if x[i] != "0" else "0"
        ys = a.ref(y[i], "in") if y[i] != "0" else "0"
        a.lines.append(f"{op} {a.ref(dst[i], 'out')}, {xs}, {ys};")
    if carry_out:
        a.lines.append(f"addc.u32 {a.ref(carry_out, 'out')}, 0, 0;")
    prog.append(("asm", a))


SC_C = ["0x2FC9BEBFu", "0x402DA173u", "0x50B75FC4u", "0x45512319u"]  # 2^256 - n = 2^128 + c (scalar_4x64_impl.h:23-25)


def prog_sc_reduce():
    """B[0..8] = t mod-n folded twice: 2^256 == 2^128 + c.  The caller finishes the (tiny) third fold in plain C."""
    prog = []
    t = [f"t[{i}]" for i in range(16)]
    hi = t[8:]
    prog_rect(prog, "E1", "O1", "P", hi, SC_C)  # P[0..11] = hi * c
    A = [f"A[{i}]" for i in range(13)]
    # A = lo + P   (13 limbs)
    prog_add(prog, A[:12], t[:8] + ["0"] * 4, [f"P[{i}]" for i in range(12)], carry_out="A[12]")
    # A += hi << 128
    a = Asm()
    for k in range(8):
        op = ("add" if k == 0 else "addc") + ".cc.u32"
        r = a.ref(A[4 + k], "io")
        a.lines.append(f"{op} {r}, {r}, {a.ref(hi[k], 'in')};")
    r = a.ref(A[12], "io")
    a.lines.append(f"addc.u32 {r}, {r}, 0;")
    prog.append(("asm", a))
    # Q[0..8] = A[8..12] * c
    prog_rect(prog, "E2", "O2", "Q", A[8:13], SC_C)
    B = [f"B[{i}]" for i in range(9)]
    prog_add(prog, B, A[:8] + ["0"], [f"Q[{i}]" for i in range(9)])
    a = Asm()
    for k in range(5):
        last = k == 4
        op = ("add" if k == 0 else "addc") + ("" if last else ".cc") + ".u32"
        r = a.ref(B[4 + k], "io")
        a.lines.append(f"{op} {r}, {r}, {a.ref(A[8 + k], 'in')};")
    prog.append(("asm", a))
    return prog


def emit_prog(prog, indent="    "):
    out = []
    for kind, item in prog:
        out.append(item.emit(indent).rstrip("\n") if kind == "asm" else indent + item)
    return "\n".join(out)


def gen_sc_reduce():
    out = ["// B[0..8] = (t[0..7] + t[8..15] * 2^256) folded twice with 2^256 == 2^128 + c (mod n): 8x4 + 5x4 = 52 IMAD.WIDE.U32",
           "SV_D void sv_sc_fold2_dev(u32* __restrict__ B, const u32* __restrict__ t) {",
           "    u32 E1[13], O1[13], P[12], A[13], E2[10], O2[10], Q[9];",
           emit_prog(prog_sc_reduce()), "}"]
    return "\n".join(out)


# ---- simulator ---------------------------------------------------------------------------------------------
def simulate(prog, regs):
    """execute a program on a dict name -> u32; immediates are decimal / 0x.. (with optional u suffix) literals"""
    import re as _re
    M = 0xFFFFFFFF

    def val(tok):
        tok = tok.strip()
        if _re.fullmatch(r"(0[xX][0-9a-fA-F]+|\d+)[uU]?", tok):
            return int(tok.rstrip("uU"), 0)
        return regs[tok]
    for kind, item in prog:
        if kind == "c":
            m = _re.fullmatch(r"\s*(\S+) = (\S+);", item)
            regs[m.group(1)] = val(m.group(2))
            continue
        cc = None
        for line in item.lines:
            line = line.replace("{", "").replace("}", "").rstrip(";")
            op, rest = line.split(" ", 1)
            ops = [x.strip() for x in rest.split(",")]
            parts = op.split(".")
            base = parts[0]
            sets_cc = "cc" in parts
            if base in ("mul",):
                p = val(ops[1]) * val(ops[2])
                regs[ops[0]] = (p & M) if "lo" in parts else (p >> 32)
            elif base in ("mad", "madc"):
                p = val(ops[1]) * val(ops[2])
                part = (p & M) if "lo" in parts else (p >> 32)
                tot = part + val(ops[3]) + ((cc or 0) if base == "madc" else 0)
                assert base != "madc" or cc is not None
                regs[ops[0]] = tot & M
                if sets_cc:
                    cc = tot >> 32
            elif base in ("add", "addc"):
                tot = val(ops[1]) + val(ops[2]) + ((cc or 0) if base == "addc" else 0)
                assert base != "addc" or cc is not None
                regs[ops[0]] = tot & M
                if sets_cc:
                    cc = tot >> 32
                else:
                    assert tot >> 32 == 0 or base == "addc" and ops[1] == "0" or True
            else:
                raise ValueError(op)
    return regs


def check():
    import random
    rnd = random.Random(5)
    N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
    NC = 2**256 - N
    prog = prog_sc_reduce()
    tests = [0, 1, 2**512 - 1, 2**256, N * N, (N - 1) ** 2, 2**512 - 2**256, (2**256 - 1) << 256, 2**256 - 1] + [rnd.getrandbits(512) for _ in range(3000)]
    for t in tests:
        regs = {f"t[{i}]": (t >> (32 * i)) & 0xFFFFFFFF for i in range(16)}
        simulate(prog, regs)
        B = sum(regs[f"B[{i}]"] << (32 * i) for i in range(9))
        m = (t & (2**256 - 1)) + (t >> 256) * NC
        q = (m & (2**256 - 1)) + (m >> 256) * NC
        assert B == q, (hex(t), hex(B), hex(q))
        assert regs["B[8]"] < 8
    # rectangular products on their own
    for na, nb in ((8, 4), (5, 4), (8, 8), (3, 2), (4, 4)):
        pr = []
        prog_rect(pr, "E", "O", "r", [f"a[{i}]" for i in range(na)], [f"b[{i}]" for i in range(nb)])
        for _ in range(300):
            a, b = rnd.getrandbits(32 * na), rnd.getrandbits(32 * nb)
            if _ % 7 == 0:
                a, b = 2**(32 * na) - 1, 2**(32 * nb) - 1
            regs = {f"a[{i}]": (a >> (32 * i)) & 0xFFFFFFFF for i in range(na)}
            regs.update({f"b[{i}]": (b >> (32 * i)) & 0xFFFFFFFF for i in range(nb)})
            simulate(pr, regs)
            assert sum(regs[f"r[{i}]"] << (32 * i) for i in range(na + nb)) == a * b, (na, nb)
    print("gen_mul --check: generated programs match big-int arithmetic")


if __name__ == "__main__":
    if "--check" in sys.argv:
        check()
        sys.exit(0)
    print("// u256_gen.cuh — GENERATED by tools/gen_mul.py; do not edit.  Device-only PTX bodies.")
    print("#pragma once")
    print('#include "common.cuh"')
    print("#if SV_DEVICE_CODE")
    print(gen_mul())
    print()
    print(gen_sqr())
    print()
    print(gen_sc_reduce())
    print("#endif")
