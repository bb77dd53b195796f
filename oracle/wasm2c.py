#!/usr/bin/env python3
"""wasm2c.py -- TEST INFRASTRUCTURE (oracle), not product code.

Translate the reference's *shipped* WebAssembly binaries
(/root/reference/wasmcurves/build/bls12381.wasm, bn128.wasm -- builds of the
upstream wasmcurves algorithm, wasmcurves/src/build_multiexp.js:20-472) into
plain C so that the reference's own g1m_multiexpAffine / _chunk / g1m_normalize
/ f1m_* run natively on the CPU.  Output goes to oracle/_ref/ (git-ignored);
nothing of the reference is copied into the repository history.

The modules only use the integer MVP subset: one imported memory, no tables,
no globals, no floats, no memory.grow.  The translator
is generic over that subset and fails loudly on anything else.

Every export `name` becomes a C symbol `name` with uint32_t/uint64_t params;
additionally:
    int      wasm_init(uint32_t pages);   // allocates linear memory, copies data segments
    uint8_t* wasm_mem(void);              // base of linear memory
    uint32_t wasm_mem_bytes(void);
"""
import sys, struct

I32, I64 = 0x7F, 0x7E
CT = {I32: "uint32_t", I64: "uint64_t"}


class Reader:
    def __init__(self, b, pos=0, end=None):
        self.b, self.pos, self.end = b, pos, len(b) if end is None else end

    def byte(self):
        v = self.b[self.pos]; self.pos += 1; return v

    def u(self):  # unsigned LEB128
        r = s = 0
        while True:
            x = self.byte(); r |= (x & 0x7F) << s; s += 7
            if not x & 0x80: return r

    def s(self, bits):  # signed LEB128
        r = s = 0
        while True:
            x = self.byte(); r |= (x & 0x7F) << s; s += 7
            if not x & 0x80:
                if x & 0x40: r -= 1 << s
                return r

    def name(self):
        n = self.u(); v = self.b[self.pos:self.pos + n]; self.pos += n; return v.decode()

    def eof(self): return self.pos >= self.end


def parse(b):
    assert b[:8] == b"\0asm\x01\0\0\0", "not a wasm v1 module"
    r = Reader(b, 8)
    m = dict(types=[], imports=[], funcs=[], exports=[], code=[], data=[], mem=None)
    while not r.eof():
        sid = r.byte(); size = r.u(); end = r.pos + size
        s = Reader(b, r.pos, end)
        if sid == 1:
            for _ in range(s.u()):
                assert s.byte() == 0x60
                ps = [s.byte() for _ in range(s.u())]
                rs = [s.byte() for _ in range(s.u())]
                m["types"].append((ps, rs))
        elif sid == 2:
            for _ in range(s.u()):
                mod, fld, kind = s.name(), s.name(), s.byte()
                assert kind == 2, "only a memory import is supported"
                flags = s.u(); mn = s.u(); mx = s.u() if flags & 1 else None
                m["mem"] = (mn, mx); m["imports"].append((mod, fld))
        elif sid == 3:
            m["funcs"] = [s.u() for _ in range(s.u())]
        elif sid == 5:
            for _ in range(s.u()):
                flags = s.u(); mn = s.u(); mx = s.u() if flags & 1 else None
                m["mem"] = (mn, mx)
        elif sid == 7:
            for _ in range(s.u()):
                nm, kind, idx = s.name(), s.byte(), s.u()
                m["exports"].append((nm, kind, idx))
        elif sid == 10:
            for _ in range(s.u()):
                sz = s.u(); fend = s.pos + sz
                f = Reader(b, s.pos, fend)
                locs = []
                for _ in range(f.u()):
                    n = f.u(); t = f.byte(); locs += [t] * n
                m["code"].append((locs, f.pos, fend))
                s.pos = fend
        elif sid == 11:
            for _ in range(s.u()):
                assert s.u() == 0
                assert s.byte() == 0x41; off = s.s(32); assert s.byte() == 0x0B
                n = s.u(); m["data"].append((off & 0xFFFFFFFF, b[s.pos:s.pos + n])); s.pos += n
        elif sid in (0,):
            pass
        elif sid in (4, 6, 8, 9, 12):
            if size and sid in (4, 6, 8, 9):
                cnt = Reader(b, r.pos, end).u()
                assert cnt == 0, f"unsupported section {sid} with {cnt} entries"
        r.pos = end
    return m


LOADS = {  # opcode: (result type, nbytes, signed, C source type)
    0x28: (I32, 4, False), 0x29: (I64, 8, False),
    0x2C: (I32, 1, True), 0x2D: (I32, 1, False), 0x2E: (I32, 2, True), 0x2F: (I32, 2, False),
    0x30: (I64, 1, True), 0x31: (I64, 1, False), 0x32: (I64, 2, True), 0x33: (I64, 2, False),
    0x34: (I64, 4, True), 0x35: (I64, 4, False),
}
STORES = {0x36: (I32, 4), 0x37: (I64, 8), 0x3A: (I32, 1), 0x3B: (I32, 2), 0x3C: (I64, 1), 0x3D: (I64, 2), 0x3E: (I64, 4)}
UT = {1: "uint8_t", 2: "uint16_t", 4: "uint32_t", 8: "uint64_t"}
ST = {1: "int8_t", 2: "int16_t", 4: "int32_t", 8: "int64_t"}

CMP32 = {0x46: ("==", 0), 0x47: ("!=", 0), 0x48: ("<", 1), 0x49: ("<", 0), 0x4A: (">", 1), 0x4B: (">", 0),
         0x4C: ("<=", 1), 0x4D: ("<=", 0), 0x4E: (">=", 1), 0x4F: (">=", 0)}
CMP64 = {k + 0x0B: v for k, v in CMP32.items()}
BIN32 = {0x6A: "+", 0x6B: "-", 0x6C: "*", 0x71: "&", 0x72: "|", 0x73: "^"}
BIN64 = {0x7C: "+", 0x7D: "-", 0x7E: "*", 0x83: "&", 0x84: "|", 0x85: "^"}


class FuncGen:
    def __init__(self, mod, fidx, fname_of):
        self.m, self.fidx, self.fname_of = mod, fidx, fname_of
        ps, rs = mod["types"][mod["funcs"][fidx]]
        locs, self.start, self.end = mod["code"][fidx]
        self.ps, self.rs = ps, rs
        self.ltypes = ps + locs
        self.out = []
        self.maxdepth = 0
        self.nlabel = 0

    def v(self, d, t):
        self.maxdepth = max(self.maxdepth, d + 1)
        return ("si%d" if t == I32 else "sj%d") % d

    def emit(self, s, ind):
        self.out.append("  " * ind + s)

    def gen(self):
        r = Reader(self.m["b"], self.start, self.end)
        st = []       # value stack: types
        # control stack entries: dict(kind, label, height, res, unreachable, used)
        ctl = [dict(kind="func", label=self.newlabel(), height=0, res=self.rs, used=False)]
        unreachable = False
        skip_depth = 0
        ind = 1

        def push(t):
            st.append(t); return self.v(len(st) - 1, t)

        def pop(t=None):
            tt = st.pop()
            if t is not None: assert tt == t, (hex(op), tt, t, self.fidx)
            return self.v(len(st), tt)

        def branch_to(n):
            """code to branch to the n-th enclosing label (moving a result value if needed)"""
            c = ctl[-1 - n]
            c["used"] = True
            code = ""
            if c["kind"] == "func":
                if c["res"]:
                    return "return %s;" % self.v(len(st) - 1, st[-1])
                return "return;"
            if c["kind"] != "loop" and c["res"]:
                t = c["res"][0]
                src = self.v(len(st) - 1, t); dst = self.v(c["height"], t)
                if src != dst: code = "%s = %s; " % (dst, src)
            return code + "goto L%d;" % c["label"]

        while not r.eof():
            op = r.byte()
            if unreachable:
                # skip until matching end/else of the current frame
                if op in (0x02, 0x03, 0x04):
                    r.byte() if True else None; skip_depth += 1; continue
                if op == 0x0B:
                    if skip_depth: skip_depth -= 1; continue
                elif op == 0x05:
                    if skip_depth: continue
                else:
                    self.skip_imm(r, op); continue
            if op == 0x00:
                self.emit("__builtin_trap();", ind); unreachable = True; skip_depth = 0
            elif op == 0x01:
                pass
            elif op in (0x02, 0x03, 0x04):
                bt = r.byte()
                res = [] if bt == 0x40 else [bt]
                assert bt in (0x40, I32, I64)
                if op == 0x04:
                    cond = pop(I32)
                lab = self.newlabel()
                kind = {0x02: "block", 0x03: "loop", 0x04: "if"}[op]
                ctl.append(dict(kind=kind, label=lab, height=len(st), res=res, used=False, has_else=False))
                if op == 0x02:
                    self.emit("{", ind)
                elif op == 0x03:
                    self.emit("L%d: ; {" % lab, ind)
                else:
                    self.emit("if (%s) {" % cond, ind)
                ind += 1
            elif op == 0x05:  # else
                c = ctl[-1]; assert c["kind"] == "if"
                if not unreachable and c["res"]:
                    t = c["res"][0]; src = pop(t); dst = self.v(c["height"], t)
                    if src != dst: self.emit("%s = %s;" % (dst, src), ind)
                del st[c["height"]:]
                unreachable = False
                c["has_else"] = True
                self.emit("} else {", ind - 1)
            elif op == 0x0B:  # end
                c = ctl.pop()
                if c["kind"] == "func":
                    if not unreachable:
                        if c["res"]:
                            self.emit("return %s;" % pop(c["res"][0]), ind)
                    elif c["res"]:
                        pass
                    break
                if not unreachable and c["res"]:
                    t = c["res"][0]; src = pop(t); dst = self.v(c["height"], t)
                    if src != dst: self.emit("%s = %s;" % (dst, src), ind)
                del st[c["height"]:]
                for t in c["res"]: st.append(t); self.v(len(st) - 1, t)
                ind -= 1
                if c["kind"] == "loop":
                    self.emit("}", ind)
                else:
                    self.emit("}", ind)
                    if c["used"]: self.emit("L%d: ;" % c["label"], ind)
                unreachable = False
            elif op == 0x0C:
                n = r.u(); self.emit(branch_to(n), ind); unreachable = True; skip_depth = 0
            elif op == 0x0D:
                n = r.u(); cond = pop(I32)
                self.emit("if (%s) { %s }" % (cond, branch_to(n)), ind)
            elif op == 0x0F:
                self.emit(branch_to(len(ctl) - 1), ind); unreachable = True; skip_depth = 0
            elif op == 0x10:
                fi = r.u()
                ps, rs = self.m["types"][self.m["funcs"][fi]]
                args = [pop(t) for t in reversed(ps)][::-1]
                call = "%s(%s)" % (self.fname_of(fi), ", ".join(args))
                if rs: self.emit("%s = %s;" % (push(rs[0]), call), ind)
                else: self.emit(call + ";", ind)
            elif op == 0x1A:
                pop()
            elif op == 0x1B:
                c = pop(I32); t = st[-1]; b_ = pop(t); a_ = pop(t)
                self.emit("%s = %s ? %s : %s;" % (push(t), c, a_, b_), ind)
            elif op == 0x20:
                i = r.u(); t = self.ltypes[i]; self.emit("%s = l%d;" % (push(t), i), ind)
            elif op == 0x21:
                i = r.u(); t = self.ltypes[i]; self.emit("l%d = %s;" % (i, pop(t)), ind)
            elif op == 0x22:
                i = r.u(); t = self.ltypes[i]; self.emit("l%d = %s;" % (i, self.v(len(st) - 1, t)), ind); assert st[-1] == t
            elif op in LOADS:
                t, nb, sg = LOADS[op]; r.u(); off = r.u()
                a = pop(I32)
                ld = "ld%d((uint64_t)%s + %du)" % (nb * 8, a, off)
                if sg: ld = "(%s)(%s)(%s)%s" % (CT[t], "int32_t" if t == I32 else "int64_t", ST[nb], ld)
                self.emit("%s = %s;" % (push(t), ld), ind)
            elif op in STORES:
                t, nb = STORES[op]; r.u(); off = r.u()
                val = pop(t); a = pop(I32)
                self.emit("st%d((uint64_t)%s + %du, (%s)%s);" % (nb * 8, a, off, UT[nb], val), ind)
            elif op == 0x41:
                c = r.s(32) & 0xFFFFFFFF; self.emit("%s = %du;" % (push(I32), c), ind)
            elif op == 0x42:
                c = r.s(64) & 0xFFFFFFFFFFFFFFFF; self.emit("%s = %dull;" % (push(I64), c), ind)
            elif op == 0x45:
                a = pop(I32); self.emit("%s = (%s == 0);" % (push(I32), a), ind)
            elif op == 0x50:
                a = pop(I64); self.emit("%s = (%s == 0);" % (push(I32), a), ind)
            elif op in CMP32 or op in CMP64:
                t = I32 if op in CMP32 else I64
                o, sg = (CMP32 if op in CMP32 else CMP64)[op]
                b_ = pop(t); a_ = pop(t)
                cast = ("(int32_t)" if t == I32 else "(int64_t)") if sg else ""
                self.emit("%s = (%s%s %s %s%s);" % (push(I32), cast, a_, o, cast, b_), ind)
            elif op in (0x67, 0x68, 0x69):
                a = pop(I32)
                e = {0x67: "(%s ? (uint32_t)__builtin_clz(%s) : 32u)", 0x68: "(%s ? (uint32_t)__builtin_ctz(%s) : 32u)",
                     0x69: "(uint32_t)__builtin_popcount(%s)"}[op]
                self.emit("%s = %s;" % (push(I32), e.replace("%s", a)), ind)
            elif op in (0x79, 0x7A, 0x7B):
                a = pop(I64)
                e = {0x79: "(%s ? (uint64_t)__builtin_clzll(%s) : 64ull)", 0x7A: "(%s ? (uint64_t)__builtin_ctzll(%s) : 64ull)",
                     0x7B: "(uint64_t)__builtin_popcountll(%s)"}[op]
                self.emit("%s = %s;" % (push(I64), e.replace("%s", a)), ind)
            elif op in BIN32 or op in BIN64:
                t = I32 if op in BIN32 else I64
                o = (BIN32 if op in BIN32 else BIN64)[op]
                b_ = pop(t); a_ = pop(t)
                self.emit("%s = %s %s %s;" % (push(t), a_, o, b_), ind)
            elif op in (0x6D, 0x6E, 0x6F, 0x70, 0x7F, 0x80, 0x81, 0x82):
                t = I32 if op < 0x7C else I64
                k = (op - 0x6D) if t == I32 else (op - 0x7F)
                b_ = pop(t); a_ = pop(t)
                sc = "(int32_t)" if t == I32 else "(int64_t)"
                o = "/" if k < 2 else "%"
                self.emit("if (%s == 0) __builtin_trap();" % b_, ind)
                if k in (0, 2):  # signed
                    if k == 0:
                        e = "(%s)(%s%s / %s%s)" % (CT[t], sc, a_, sc, b_)
                    else:
                        e = "((%s%s == -1) ? 0 : (%s)(%s%s %% %s%s))" % (sc, b_, CT[t], sc, a_, sc, b_)
                else:
                    e = "%s %s %s" % (a_, o, b_)
                self.emit("%s = %s;" % (push(t), e), ind)
            elif op in (0x74, 0x75, 0x76, 0x77, 0x78, 0x86, 0x87, 0x88, 0x89, 0x8A):
                t = I32 if op < 0x79 else I64
                k = (op - 0x74) if t == I32 else (op - 0x86)
                bits = 32 if t == I32 else 64
                b_ = pop(t); a_ = pop(t)
                sh = "(%s & %d)" % (b_, bits - 1)
                sc = "(int32_t)" if t == I32 else "(int64_t)"
                if k == 0: e = "%s << %s" % (a_, sh)
                elif k == 1: e = "(%s)(%s%s >> %s)" % (CT[t], sc, a_, sh)
                elif k == 2: e = "%s >> %s" % (a_, sh)
                elif k == 3: e = "(%s << %s) | (%s >> ((%d - %s) & %d))" % (a_, sh, a_, bits, sh, bits - 1)
                else: e = "(%s >> %s) | (%s << ((%d - %s) & %d))" % (a_, sh, a_, bits, sh, bits - 1)
                self.emit("%s = %s;" % (push(t), e), ind)
            elif op == 0xA7:
                a = pop(I64); self.emit("%s = (uint32_t)%s;" % (push(I32), a), ind)
            elif op == 0xAC:
                a = pop(I32); self.emit("%s = (uint64_t)(int64_t)(int32_t)%s;" % (push(I64), a), ind)
            elif op == 0xAD:
                a = pop(I32); self.emit("%s = (uint64_t)%s;" % (push(I64), a), ind)
            else:
                raise NotImplementedError("opcode 0x%02x in function %d" % (op, self.fidx))
        # assemble
        name = self.fname_of(self.fidx)
        rt = CT[self.rs[0]] if self.rs else "void"
        params = ", ".join("%s l%d" % (CT[t], i) for i, t in enumerate(self.ps)) or "void"
        hdr = ["%s %s(%s) {" % (rt, name, params)]
        for i in range(len(self.ps), len(self.ltypes)):
            hdr.append("  %s l%d = 0;" % (CT[self.ltypes[i]], i))
        if self.maxdepth:
            hdr.append("  uint32_t " + ", ".join("si%d = 0" % d for d in range(self.maxdepth)) + ";")
            hdr.append("  uint64_t " + ", ".join("sj%d = 0" % d for d in range(self.maxdepth)) + ";")
            hdr.append("  " + " ".join("(void)si%d; (void)sj%d;" % (d, d) for d in range(self.maxdepth)))
        tail = []
        if self.rs and unreachable:
            tail.append("  __builtin_unreachable();")
        return "\n".join(hdr + self.out + tail + ["}"])

    def newlabel(self):
        self.nlabel += 1; return self.nlabel

    def skip_imm(self, r, op):
        if op in (0x0C, 0x0D, 0x10, 0x20, 0x21, 0x22): r.u()
        elif op in LOADS or op in STORES: r.u(); r.u()
        elif op == 0x41: r.s(32)
        elif op == 0x42: r.s(64)
        elif op == 0x0E:
            for _ in range(r.u() + 1): r.u()
        elif op == 0x11: r.u(); r.u()
        elif op in (0x3F, 0x40): r.byte()


def translate(wasm_bytes, modname):
    m = parse(wasm_bytes); m["b"] = wasm_bytes
    expname = {}
    for nm, kind, idx in m["exports"]:
        if kind == 0: expname.setdefault(idx, nm)
    nimp = 0  # function imports unsupported (asserted in parse)

    def fname_of(i): return "f%d" % i

    out = ["/* GENERATED by oracle/wasm2c.py from the reference's %s.wasm -- do not commit */" % modname,
           "#include <stdint.h>", "#include <string.h>", "#include <stdlib.h>",
           "static uint8_t* mem; static uint32_t mem_bytes;",
           "#define LD(T) T v; memcpy(&v, mem + a, sizeof v); return v;",
           "static inline uint8_t  ld8 (uint64_t a){ LD(uint8_t) }",
           "static inline uint16_t ld16(uint64_t a){ LD(uint16_t) }",
           "static inline uint32_t ld32(uint64_t a){ LD(uint32_t) }",
           "static inline uint64_t ld64(uint64_t a){ LD(uint64_t) }",
           "static inline void st8 (uint64_t a, uint8_t  v){ memcpy(mem + a, &v, sizeof v); }",
           "static inline void st16(uint64_t a, uint16_t v){ memcpy(mem + a, &v, sizeof v); }",
           "static inline void st32(uint64_t a, uint32_t v){ memcpy(mem + a, &v, sizeof v); }",
           "static inline void st64(uint64_t a, uint64_t v){ memcpy(mem + a, &v, sizeof v); }"]
    # prototypes
    for i, ti in enumerate(m["funcs"]):
        ps, rs = m["types"][ti]
        out.append("static %s f%d(%s);" % (CT[rs[0]] if rs else "void", i, ", ".join(CT[t] for t in ps) or "void"))
    for i in range(len(m["funcs"])):
        out.append("static " + FuncGen(m, i, fname_of).gen())
    # data + init
    out.append("static const struct { uint32_t off, len; const uint8_t* p; } segs[] = {")
    blobs = []
    for k, (off, data) in enumerate(m["data"]):
        blobs.append("static const uint8_t seg%d[] = {%s};" % (k, ",".join(str(x) for x in data) or "0"))
        out.append("  {%du, %du, seg%d}," % (off, len(data), k))
    out.append("};")
    out[out.index("static const struct { uint32_t off, len; const uint8_t* p; } segs[] = {"):0] = blobs
    out.append("""
int wasm_init(uint32_t pages) {
  if (pages < %du) pages = %du;
  uint64_t nb = (uint64_t)pages << 16; if (nb > 0xFFFF0000ull) nb = 0xFFFF0000ull;
  free(mem); mem = (uint8_t*)calloc(nb + 16, 1); if (!mem) return -1;
  mem_bytes = (uint32_t)nb;
  for (unsigned i = 0; i < sizeof segs / sizeof segs[0]; i++) memcpy(mem + segs[i].off, segs[i].p, segs[i].len);
  return 0;
}
uint8_t* wasm_mem(void) { return mem; }
uint32_t wasm_mem_bytes(void) { return mem_bytes; }""" % (m["mem"][0], m["mem"][0]))
    for nm, kind, idx in m["exports"]:
        if kind != 0: continue
        ps, rs = m["types"][m["funcs"][idx]]
        rt = CT[rs[0]] if rs else "void"
        params = ", ".join("%s a%d" % (CT[t], i) for i, t in enumerate(ps)) or "void"
        args = ", ".join("a%d" % i for i in range(len(ps)))
        out.append("%s %s(%s) { %sf%d(%s); }" % (rt, nm, params, "return " if rs else "", idx, args))
    return "\n".join(out) + "\n", m


if __name__ == "__main__":
    src, dst, modname = sys.argv[1], sys.argv[2], sys.argv[3]
    code, m = translate(open(src, "rb").read(), modname)
    open(dst, "w").write(code)
    print("%s: %d functions, %d exports, %d data segments, mem min %d pages -> %s (%d lines)" % (
        src, len(m["funcs"]), len(m["exports"]), len(m["data"]), m["mem"][0], dst, code.count("\n")))
