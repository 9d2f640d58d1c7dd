"""JVM class-file reader (constant pool, fields, methods, Code attributes) and disassembler -- TEST INFRASTRUCTURE.

The reference is Java and the image has no JVM; its arithmetic containers (librec.data.DenseMatrix / DenseVector /
SparseMatrix) ship as bytecode only (lib/librec-v1.4-alpha.jar).  This module lets the tests READ that bytecode --
and, through tests/tools/minijvm.py, EXECUTE it -- so the CPU oracle is pinned to an artefact the reference holds
instead of to a line-by-line reading of the sources.  Format: JVMS chapter 4 (class file), chapter 6 (instructions).
"""
from __future__ import annotations

import struct
import zipfile
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

# mnemonic, operand format.  '' none; 'b' s1; 'B' u1; 's' s2; 'S' u2 (cp index); 'Bb' u1 s1 (iinc); 'j' s2 branch;
# 'J' s4 branch; 'SBB' invokeinterface; 'SS0' invokedynamic; 'SB' multianewarray; 'T' tableswitch; 'L' lookupswitch; 'W' wide
_OPS: Dict[int, Tuple[str, str]] = {}


def _fill():
    names0 = ("nop aconst_null iconst_m1 iconst_0 iconst_1 iconst_2 iconst_3 iconst_4 iconst_5 lconst_0 lconst_1 "
              "fconst_0 fconst_1 fconst_2 dconst_0 dconst_1").split()
    for i, n in enumerate(names0):
        _OPS[i] = (n, "")
    _OPS[16] = ("bipush", "b")
    _OPS[17] = ("sipush", "s")
    _OPS[18] = ("ldc", "B")
    _OPS[19] = ("ldc_w", "S")
    _OPS[20] = ("ldc2_w", "S")
    for i, n in enumerate("iload lload fload dload aload".split()):
        _OPS[21 + i] = (n, "B")
    k = 26
    for t in "ilfda":
        for i in range(4):
            _OPS[k] = (f"{t}load_{i}", "")
            k += 1
    for i, n in enumerate("iaload laload faload daload aaload baload caload saload".split()):
        _OPS[46 + i] = (n, "")
    for i, n in enumerate("istore lstore fstore dstore astore".split()):
        _OPS[54 + i] = (n, "B")
    k = 59
    for t in "ilfda":
        for i in range(4):
            _OPS[k] = (f"{t}store_{i}", "")
            k += 1
    for i, n in enumerate("iastore lastore fastore dastore aastore bastore castore sastore".split()):
        _OPS[79 + i] = (n, "")
    for i, n in enumerate("pop pop2 dup dup_x1 dup_x2 dup2 dup2_x1 dup2_x2 swap".split()):
        _OPS[87 + i] = (n, "")
    k = 96
    for op in "add sub mul div rem neg".split():
        for t in "ilfd":
            _OPS[k] = (t + op, "")
            k += 1
    for i, n in enumerate("ishl lshl ishr lshr iushr lushr iand land ior lor ixor lxor".split()):
        _OPS[120 + i] = (n, "")
    _OPS[132] = ("iinc", "Bb")
    for i, n in enumerate("i2l i2f i2d l2i l2f l2d f2i f2l f2d d2i d2l d2f i2b i2c i2s".split()):
        _OPS[133 + i] = (n, "")
    for i, n in enumerate("lcmp fcmpl fcmpg dcmpl dcmpg".split()):
        _OPS[148 + i] = (n, "")
    for i, n in enumerate("ifeq ifne iflt ifge ifgt ifle if_icmpeq if_icmpne if_icmplt if_icmpge if_icmpgt if_icmple "
                          "if_acmpeq if_acmpne goto jsr".split()):
        _OPS[153 + i] = (n, "j")
    _OPS[169] = ("ret", "B")
    _OPS[170] = ("tableswitch", "T")
    _OPS[171] = ("lookupswitch", "L")
    for i, n in enumerate("ireturn lreturn freturn dreturn areturn return".split()):
        _OPS[172 + i] = (n, "")
    for i, n in enumerate("getstatic putstatic getfield putfield invokevirtual invokespecial invokestatic".split()):
        _OPS[178 + i] = (n, "S")
    _OPS[185] = ("invokeinterface", "SBB")
    _OPS[186] = ("invokedynamic", "SS0")
    _OPS[187] = ("new", "S")
    _OPS[188] = ("newarray", "B")
    _OPS[189] = ("anewarray", "S")
    _OPS[190] = ("arraylength", "")
    _OPS[191] = ("athrow", "")
    _OPS[192] = ("checkcast", "S")
    _OPS[193] = ("instanceof", "S")
    _OPS[194] = ("monitorenter", "")
    _OPS[195] = ("monitorexit", "")
    _OPS[196] = ("wide", "W")
    _OPS[197] = ("multianewarray", "SB")
    _OPS[198] = ("ifnull", "j")
    _OPS[199] = ("ifnonnull", "j")
    _OPS[200] = ("goto_w", "J")
    _OPS[201] = ("jsr_w", "J")


_fill()


@dataclass
class Insn:
    pc: int
    op: str
    args: tuple = ()      # decoded operands (ints; branch targets are absolute pcs)
    ref: object = None    # resolved constant-pool operand: ("Class", name) / (class, name, descriptor) / constant value

    def __repr__(self):
        r = "" if self.ref is None else f" {self.ref}"
        a = "" if not self.args or self.ref is not None else " " + " ".join(map(str, self.args))
        return f"{self.pc:4d} {self.op}{a}{r}"


@dataclass
class Method:
    name: str
    desc: str
    access: int
    max_stack: int = 0
    max_locals: int = 0
    code: List[Insn] = field(default_factory=list)
    exceptions: list = field(default_factory=list)   # (start_pc, end_pc, handler_pc, catch_type)
    lines: Dict[int, int] = field(default_factory=dict)  # pc -> source line

    @property
    def is_static(self):
        return bool(self.access & 0x0008)


class ClassFile:
    def __init__(self, data: bytes):
        self.d = data
        self.p = 0
        if self._u4() != 0xCAFEBABE:
            raise ValueError("not a class file")
        self.minor, self.major = self._u2(), self._u2()
        self.cp: List[object] = [None]
        n = self._u2()
        i = 1
        while i < n:
            tag = self._u1()
            if tag == 1:
                ln = self._u2()
                self.cp.append(("Utf8", self.d[self.p:self.p + ln].decode("utf-8", "replace")))
                self.p += ln
            elif tag == 3:
                self.cp.append(("Integer", struct.unpack(">i", self._take(4))[0]))
            elif tag == 4:
                self.cp.append(("Float", struct.unpack(">f", self._take(4))[0]))
            elif tag == 5:
                self.cp.append(("Long", struct.unpack(">q", self._take(8))[0]))
                self.cp.append(None)
                i += 1
            elif tag == 6:
                self.cp.append(("Double", struct.unpack(">d", self._take(8))[0]))
                self.cp.append(None)
                i += 1
            elif tag in (7, 8, 16, 19, 20):
                self.cp.append(({7: "Class", 8: "String", 16: "MethodType", 19: "Module", 20: "Package"}[tag], self._u2()))
            elif tag in (9, 10, 11, 12, 17, 18):
                self.cp.append(({9: "Fieldref", 10: "Methodref", 11: "InterfaceMethodref", 12: "NameAndType", 17: "Dynamic",
                                 18: "InvokeDynamic"}[tag], self._u2(), self._u2()))
            elif tag == 15:
                self.cp.append(("MethodHandle", self._u1(), self._u2()))
            else:
                raise ValueError(f"constant pool tag {tag}")
            i += 1
        self.access = self._u2()
        self.name = self.class_name(self._u2())
        sup = self._u2()
        self.super_name = self.class_name(sup) if sup else None
        self.interfaces = [self.class_name(self._u2()) for _ in range(self._u2())]
        self.fields: Dict[str, Tuple[str, int]] = {}
        for _ in range(self._u2()):
            acc, nm, ds = self._u2(), self.utf8(self._u2()), self.utf8(self._u2())
            self._skip_attributes()
            self.fields[nm] = (ds, acc)
        self.methods: Dict[Tuple[str, str], Method] = {}
        for _ in range(self._u2()):
            acc, nm, ds = self._u2(), self.utf8(self._u2()), self.utf8(self._u2())
            m = Method(nm, ds, acc)
            for _ in range(self._u2()):
                an, ln = self.utf8(self._u2()), self._u4()
                end = self.p + ln
                if an == "Code":
                    self._code(m)
                self.p = end
            self.methods[(nm, ds)] = m

    # ---- primitive readers ----
    def _take(self, n):
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    def _u1(self):
        v = self.d[self.p]
        self.p += 1
        return v

    def _u2(self):
        v = struct.unpack_from(">H", self.d, self.p)[0]
        self.p += 2
        return v

    def _u4(self):
        v = struct.unpack_from(">I", self.d, self.p)[0]
        self.p += 4
        return v

    def _skip_attributes(self):
        for _ in range(self._u2()):
            self._u2()
            n = self._u4()  # (not `self.p += self._u4()`: the augmented assignment reads self.p before the call)
            self.p += n

    # ---- constant pool ----
    def utf8(self, i):
        return self.cp[i][1]

    def class_name(self, i):
        return self.utf8(self.cp[i][1])

    def member(self, i):
        _, ci, nti = self.cp[i]
        _, ni, di = self.cp[nti]
        return (self.class_name(ci), self.utf8(ni), self.utf8(di))

    def constant(self, i):
        c = self.cp[i]
        if c[0] in ("Integer", "Float", "Long", "Double"):
            return c
        if c[0] == "String":
            return ("String", self.utf8(c[1]))
        if c[0] == "Class":
            return ("Class", self.utf8(c[1]))
        return c

    # ---- Code attribute ----
    def _code(self, m: Method):
        m.max_stack, m.max_locals = self._u2(), self._u2()
        n = self._u4()
        base = self.p
        code = self.d[base:base + n]
        self.p = base + n
        for _ in range(self._u2()):
            s, e, h, t = self._u2(), self._u2(), self._u2(), self._u2()
            m.exceptions.append((s, e, h, self.class_name(t) if t else None))
        for _ in range(self._u2()):
            an, ln = self.utf8(self._u2()), self._u4()
            end = self.p + ln
            if an == "LineNumberTable":
                for _ in range(self._u2()):
                    pc, line = self._u2(), self._u2()
                    m.lines[pc] = line
            self.p = end
        m.code = self._disassemble(code)

    def _disassemble(self, code: bytes) -> List[Insn]:
        out: List[Insn] = []
        pc = 0
        s1 = lambda o: struct.unpack_from(">b", code, o)[0]
        s2 = lambda o: struct.unpack_from(">h", code, o)[0]
        u2 = lambda o: struct.unpack_from(">H", code, o)[0]
        s4 = lambda o: struct.unpack_from(">i", code, o)[0]
        while pc < len(code):
            opc = code[pc]
            name, fmt = _OPS[opc]
            ins = Insn(pc, name)
            nxt = pc + 1
            if fmt == "b":
                ins.args = (s1(nxt),); nxt += 1
            elif fmt == "B":
                ins.args = (code[nxt],); nxt += 1
                if name == "ldc":
                    ins.ref = self.constant(code[nxt - 1])
            elif fmt == "s":
                ins.args = (s2(nxt),); nxt += 2
            elif fmt == "S":
                idx = u2(nxt); nxt += 2
                ins.args = (idx,)
                if name in ("ldc_w", "ldc2_w"):
                    ins.ref = self.constant(idx)
                elif name in ("new", "anewarray", "checkcast", "instanceof"):
                    ins.ref = ("Class", self.class_name(idx))
                else:
                    ins.ref = self.member(idx)
            elif fmt == "Bb":
                ins.args = (code[nxt], s1(nxt + 1)); nxt += 2
            elif fmt == "j":
                ins.args = (pc + s2(nxt),); nxt += 2
            elif fmt == "J":
                ins.args = (pc + s4(nxt),); nxt += 4
            elif fmt == "SBB":
                idx = u2(nxt); ins.args = (idx, code[nxt + 2]); ins.ref = self.member(idx); nxt += 4
            elif fmt == "SS0":
                ins.args = (u2(nxt),); nxt += 4
            elif fmt == "SB":
                idx = u2(nxt); ins.args = (idx, code[nxt + 2]); ins.ref = ("Class", self.class_name(idx)); nxt += 3
            elif fmt == "T":
                nxt = (nxt + 3) & ~3
                dflt, lo, hi = s4(nxt), s4(nxt + 4), s4(nxt + 8)
                nxt += 12
                tg = [pc + s4(nxt + 4 * k) for k in range(hi - lo + 1)]
                nxt += 4 * (hi - lo + 1)
                ins.args = (pc + dflt, lo, hi, tuple(tg))
            elif fmt == "L":
                nxt = (nxt + 3) & ~3
                dflt, np_ = s4(nxt), s4(nxt + 4)
                nxt += 8
                pairs = tuple((s4(nxt + 8 * k), pc + s4(nxt + 8 * k + 4)) for k in range(np_))
                nxt += 8 * np_
                ins.args = (pc + dflt, pairs)
            elif fmt == "W":
                sub = _OPS[code[nxt]][0]
                if sub == "iinc":
                    ins.op, ins.args = "iinc", (u2(nxt + 1), s2(nxt + 3)); nxt += 5
                else:
                    ins.op, ins.args = sub, (u2(nxt + 1),); nxt += 3
            out.append(ins)
            pc = nxt
        return out

    def method(self, name: str, desc: Optional[str] = None) -> Method:
        hits = [m for (n, d), m in self.methods.items() if n == name and (desc is None or d == desc)]
        if len(hits) != 1:
            raise KeyError(f"{self.name}.{name}{desc or ''}: {len(hits)} matches {[m.desc for m in hits]}")
        return hits[0]


class Jar:
    """Lazy class loader over one or more jars."""

    def __init__(self, *paths: str):
        self.zips = [zipfile.ZipFile(p) for p in paths]
        self.cache: Dict[str, ClassFile] = {}

    def has(self, name: str) -> bool:
        if name in self.cache:
            return True
        return any(name + ".class" in z.NameToInfo for z in self.zips)

    def load(self, name: str) -> ClassFile:
        if name not in self.cache:
            for z in self.zips:
                if name + ".class" in z.NameToInfo:
                    self.cache[name] = ClassFile(z.read(name + ".class"))
                    break
            else:
                raise KeyError(name)
        return self.cache[name]

    def find_method(self, cls: str, name: str, desc: str):
        """Virtual / static method resolution: walk the superclass chain inside the jars."""
        c: Optional[str] = cls
        while c is not None and self.has(c):
            cf = self.load(c)
            m = cf.methods.get((name, desc))
            if m is not None and m.code:
                return cf, m
            c = cf.super_name
        return None, None


ARITH = {"dadd", "dsub", "dmul", "ddiv", "drem", "dneg", "fadd", "fsub", "fmul", "fdiv", "f2d", "d2f", "i2d", "i2f", "d2i",
         "l2d", "iadd", "isub", "imul", "idiv"}


def arithmetic_skeleton(m: Method, keep=ARITH, calls=True) -> List[str]:
    """The method's floating-point instructions and calls, in code order -- what the oracle's statement order must match."""
    out = []
    for ins in m.code:
        if ins.op in keep:
            out.append(ins.op)
        elif calls and ins.op.startswith("invoke"):
            c, n, _ = ins.ref
            out.append(f"{c.split('/')[-1]}.{n}")
    return out


if __name__ == "__main__":
    import sys
    jar = Jar(*sys.argv[1].split(":"))
    cf = jar.load(sys.argv[2])
    for (n, d), m in cf.methods.items():
        if len(sys.argv) > 3 and n != sys.argv[3]:
            continue
        print(f"--- {cf.name}.{n}{d} (stack {m.max_stack}, locals {m.max_locals})")
        for ins in m.code:
            ln = m.lines.get(ins.pc)
            print(("%5s " % (f"L{ln}" if ln else "")) + repr(ins))
