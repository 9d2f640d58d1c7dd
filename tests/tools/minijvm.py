"""A small JVM bytecode interpreter -- TEST INFRASTRUCTURE (parity pin).

The image has no JVM, so the reference cannot be *run*; but its class files are in /root/reference/jar and
/root/reference/lib, and the hot path (buildModel / predict / isConverged / updateLRate and librec's DenseMatrix,
DenseVector) is plain arithmetic over arrays.  This interpreter executes exactly those methods FROM THE REFERENCE'S
OWN BYTECODE: every dmul / dadd / dsub / f2d is performed in the order and with the operand widths javac emitted, on
IEEE-754 doubles (Python floats) and floats (numpy.float32).  Containers that are not arithmetic (the rating iterator,
HashMap lookups of DataDAO, Guava tables, boxed Integers, logging) are supplied by the harness as "natives".

Value representation: int/short/byte/char/boolean -> Python int (wrapped to 32 bits); long -> JLong; float ->
numpy.float32; double -> Python float; references -> JObject / Python list (arrays) / str / host objects / None.
A double or long occupies ONE entry of the Python operand stack (dup2 / pop2 look at the type).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .classfile import ClassFile, Insn, Jar, Method

F32 = np.float32


class JLong(int):
    pass


class JObject:
    __slots__ = ("cls", "f")

    def __init__(self, cls: str, **fields):
        self.cls = cls
        self.f = dict(fields)

    def __repr__(self):
        return f"<{self.cls}>"


class JavaThrow(Exception):
    def __init__(self, obj):
        super().__init__(repr(obj))
        self.obj = obj


def i32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - 0x100000000 if x & 0x80000000 else x


def i64(x: int) -> JLong:
    x &= 0xFFFFFFFFFFFFFFFF
    return JLong(x - 0x10000000000000000 if x & 0x8000000000000000 else x)


def is_cat2(v) -> bool:
    return isinstance(v, JLong) or (isinstance(v, float) and not isinstance(v, F32))


def parse_desc(desc: str) -> Tuple[List[str], str]:
    assert desc[0] == "("
    i, args = 1, []
    while desc[i] != ")":
        j = i
        while desc[j] == "[":
            j += 1
        if desc[j] == "L":
            j = desc.index(";", j)
        args.append(desc[i:j + 1])
        i = j + 1
    return args, desc[i + 1:]


def default_value(desc: str):
    c = desc[0]
    if c in "IZBSC":
        return 0
    if c == "J":
        return JLong(0)
    if c == "F":
        return F32(0)
    if c == "D":
        return 0.0
    return None


class Frame:
    __slots__ = ("cf", "m", "loc", "st")

    def __init__(self, cf: ClassFile, m: Method, args: list):
        self.cf, self.m = cf, m
        self.loc = [None] * (m.max_locals + 2)
        k = 0
        for a in args:  # category-2 values take two local slots
            self.loc[k] = a
            k += 2 if is_cat2(a) else 1
        self.st: list = []


class MiniJVM:
    def __init__(self, jar: Jar, natives: Optional[Dict[tuple, Callable]] = None):
        self.jar = jar
        self.natives: Dict[tuple, Callable] = dict(natives or {})
        self.statics: Dict[Tuple[str, str], object] = {}
        self.ops_executed = 0
        self.op_counts: Dict[str, int] = {}
        self.count_ops = False
        self._pcmaps: Dict[int, Dict[int, int]] = {}
        self.natives.setdefault(("java/lang/Object", "<init>"), lambda jvm, a: None)
        self.natives.setdefault(("java/lang/Math", "abs"), lambda jvm, a: abs(a[0]))
        self.natives.setdefault(("java/lang/Math", "sqrt"), lambda jvm, a: math.sqrt(a[0]) if a[0] >= 0 else float("nan"))
        # Math.pow(x, 2.0): HotSpot's intrinsic (and fdlibm's e_pow.c for an exact product) return x * x rounded once
        self.natives.setdefault(("java/lang/Math", "pow"), lambda jvm, a: a[0] * a[0] if a[1] == 2.0 else math.pow(a[0], a[1]))
        self.natives.setdefault(("java/lang/Double", "isNaN"), lambda jvm, a: int(a[0] != a[0]))
        self.natives.setdefault(("java/lang/Double", "isInfinite"), lambda jvm, a: int(math.isinf(a[0])))
        self.natives.setdefault(("java/lang/Double", "valueOf"), lambda jvm, a: a[0])
        self.natives.setdefault(("java/lang/Double", "doubleValue"), lambda jvm, a: a[0])
        self.natives.setdefault(("java/lang/Integer", "valueOf"), lambda jvm, a: a[0])
        self.natives.setdefault(("java/lang/Integer", "intValue"), lambda jvm, a: a[0])
        self.natives.setdefault(("java/lang/Float", "valueOf"), lambda jvm, a: a[0])
        self.natives.setdefault(("happy/coding/io/Logs", "debug"), lambda jvm, a: None)
        self.natives.setdefault(("happy/coding/io/Logs", "error"), lambda jvm, a: None)
        self.natives.setdefault(("java/lang/System", "exit"), self._exit)
        self.natives.setdefault(("java/util/List", "iterator"), lambda jvm, a: HostIterator(a[0]))
        self.natives.setdefault(("java/util/List", "size"), lambda jvm, a: len(a[0]))
        self.natives.setdefault(("java/util/List", "get"), lambda jvm, a: a[0][a[1]])
        self.natives.setdefault(("java/util/Iterator", "hasNext"), lambda jvm, a: int(a[0].has_next()))
        self.natives.setdefault(("java/util/Iterator", "next"), lambda jvm, a: a[0].next())

    @staticmethod
    def _exit(jvm, a):
        raise JavaThrow(f"System.exit({a[0]})")

    # ---- statics / fields -------------------------------------------------------------------------------
    def _declaring_class(self, cls: str, field: str) -> str:
        c = cls
        while c is not None and self.jar.has(c):
            cf = self.jar.load(c)
            if field in cf.fields:
                return c
            c = cf.super_name
        return cls

    def set_static(self, cls: str, field: str, value):
        self.statics[(self._declaring_class(cls, field), field)] = value

    def get_static(self, cls: str, field: str, desc: str = "I"):
        if field == "$assertionsDisabled":
            return 1  # `java` runs with assertions disabled unless -ea is given
        key = (self._declaring_class(cls, field), field)
        if key not in self.statics:
            return default_value(desc)
        return self.statics[key]

    # ---- method lookup ----------------------------------------------------------------------------------
    def _find(self, cls: str, name: str, desc: str):
        """Walk the superclass chain from `cls`: a native registered for a class on the way wins over bytecode."""
        c: Optional[str] = cls
        while c is not None:
            nat = self.natives.get((c, name, desc)) or self.natives.get((c, name))
            if nat is not None:
                return nat, None, None
            if not self.jar.has(c):
                break
            cf = self.jar.load(c)
            m = cf.methods.get((name, desc))
            if m is not None and m.code:
                return None, cf, m
            c = cf.super_name
        return None, None, None

    def call(self, cls: str, name: str, desc: str, args: list):
        nat, cf, m = self._find(cls, name, desc)
        if nat is not None:
            return nat(self, args)
        if m is None:
            raise NotImplementedError(f"no bytecode and no native for {cls}.{name}{desc}")
        return self.run(cf, m, args)

    def call_virtual(self, obj, owner: str, name: str, desc: str, args: list):
        start = obj.cls if isinstance(obj, JObject) else (getattr(obj, "jclass", None) or owner)
        nat, cf, m = self._find(start, name, desc)
        if nat is None and m is None and start != owner:
            nat, cf, m = self._find(owner, name, desc)
        if nat is not None:
            return nat(self, [obj] + args)
        if m is None:
            raise NotImplementedError(f"no bytecode and no native for {start}.{name}{desc} (declared on {owner})")
        return self.run(cf, m, [obj] + args)

    # ---- the interpreter loop ----------------------------------------------------------------------------
    def run(self, cf: ClassFile, m: Method, args: list):
        fr = Frame(cf, m, args)
        code = m.code
        pcmap = self._pcmaps.get(id(m))
        if pcmap is None:
            pcmap = {ins.pc: k for k, ins in enumerate(code)}
            self._pcmaps[id(m)] = pcmap
        st, loc = fr.st, fr.loc
        k = 0
        count = self.count_ops
        while True:
            ins = code[k]
            op = ins.op
            self.ops_executed += 1
            if count:
                self.op_counts[op] = self.op_counts.get(op, 0) + 1
            k += 1
            # -- loads / stores / constants (most frequent first) --
            c0 = op[0]
            if op.endswith(("load_0", "load_1", "load_2", "load_3")) and len(op) == 7:
                st.append(loc[int(op[-1])])
            elif op in ("iload", "lload", "fload", "dload", "aload"):
                st.append(loc[ins.args[0]])
            elif op.endswith(("store_0", "store_1", "store_2", "store_3")) and len(op) == 8:
                loc[int(op[-1])] = st.pop()
            elif op in ("istore", "lstore", "fstore", "dstore", "astore"):
                loc[ins.args[0]] = st.pop()
            elif op == "getfield":
                o = st.pop()
                if o is None:
                    raise JavaThrow("NullPointerException getfield " + str(ins.ref))
                _, name, desc = ins.ref
                v = o.f.get(name)
                st.append(default_value(desc) if v is None and name not in o.f else v)
            elif op == "putfield":
                v = st.pop()
                o = st.pop()
                o.f[ins.ref[1]] = v
            elif op == "getstatic":
                st.append(self.get_static(ins.ref[0], ins.ref[1], ins.ref[2]))
            elif op == "putstatic":
                self.set_static(ins.ref[0], ins.ref[1], st.pop())
            elif op == "dmul":
                b = st.pop(); st[-1] = st[-1] * b
            elif op == "dadd":
                b = st.pop(); st[-1] = st[-1] + b
            elif op == "dsub":
                b = st.pop(); st[-1] = st[-1] - b
            elif op == "ddiv":
                b = st.pop(); a = st[-1]
                if b == 0.0:
                    st[-1] = float("nan") if (a == 0.0 or a != a) else math.copysign(float("inf"), a) * math.copysign(1.0, b)
                else:
                    st[-1] = a / b
            elif op == "dneg":
                st[-1] = -st[-1]
            elif op == "f2d":
                st[-1] = float(st[-1])
            elif op == "d2f":
                with np.errstate(over="ignore"):
                    st[-1] = F32(st[-1])
            elif op == "i2d":
                st[-1] = float(st[-1])
            elif op == "i2f":
                st[-1] = F32(st[-1])
            elif op == "l2d":
                st[-1] = float(int(st[-1]))
            elif op == "i2l":
                st[-1] = JLong(st[-1])
            elif op == "l2i":
                st[-1] = i32(int(st[-1]))
            elif op in ("d2i", "f2i"):
                v = float(st[-1])
                st[-1] = 0 if v != v else max(-2 ** 31, min(2 ** 31 - 1, int(v)))
            elif op == "d2l":
                v = st[-1]
                st[-1] = JLong(0 if v != v else max(-2 ** 63, min(2 ** 63 - 1, int(v))))
            elif op in ("fadd", "fsub", "fmul", "fdiv"):
                b = st.pop(); a = st[-1]
                with np.errstate(all="ignore"):
                    st[-1] = F32(a + b if op == "fadd" else a - b if op == "fsub" else a * b if op == "fmul" else a / b)
            elif op == "daload" or op == "aaload" or op == "iaload" or op == "faload" or op == "laload" or op == "baload":
                i = st.pop(); arr = st.pop()
                if arr is None:
                    raise JavaThrow("NullPointerException " + op)
                if not 0 <= i < len(arr):
                    raise JavaThrow(f"ArrayIndexOutOfBoundsException {i}")
                st.append(arr[i])
            elif op in ("dastore", "aastore", "iastore", "fastore", "lastore", "bastore"):
                v = st.pop(); i = st.pop(); arr = st.pop()
                if not 0 <= i < len(arr):
                    raise JavaThrow(f"ArrayIndexOutOfBoundsException {i}")
                arr[i] = v
            elif op == "arraylength":
                st[-1] = len(st[-1])
            elif op.startswith("iconst_"):
                st.append(-1 if op == "iconst_m1" else int(op[-1]))
            elif op in ("dconst_0", "dconst_1"):
                st.append(float(op[-1]))
            elif op in ("fconst_0", "fconst_1", "fconst_2"):
                st.append(F32(int(op[-1])))
            elif op in ("lconst_0", "lconst_1"):
                st.append(JLong(int(op[-1])))
            elif op == "aconst_null":
                st.append(None)
            elif op in ("bipush", "sipush"):
                st.append(ins.args[0])
            elif op in ("ldc", "ldc_w", "ldc2_w"):
                kind, val = ins.ref[0], ins.ref[1]
                st.append(F32(val) if kind == "Float" else JLong(val) if kind == "Long" else ins.ref if kind == "Class" else val)
            elif op == "iinc":
                loc[ins.args[0]] = i32(loc[ins.args[0]] + ins.args[1])
            elif op in ("iadd", "isub", "imul"):
                b = st.pop(); a = st[-1]
                st[-1] = i32(a + b if op == "iadd" else a - b if op == "isub" else a * b)
            elif op in ("idiv", "irem"):
                b = st.pop(); a = st[-1]
                if b == 0:
                    raise JavaThrow("ArithmeticException / by zero")
                q = abs(a) // abs(b) * (1 if (a < 0) == (b < 0) else -1)
                st[-1] = i32(q if op == "idiv" else a - q * b)
            elif op == "ineg":
                st[-1] = i32(-st[-1])
            elif op in ("ladd", "lsub", "lmul"):
                b = int(st.pop()); a = int(st[-1])
                st[-1] = i64(a + b if op == "ladd" else a - b if op == "lsub" else a * b)
            elif op in ("ishl", "ishr", "iushr", "iand", "ior", "ixor"):
                b = st.pop(); a = st[-1]
                if op == "ishl": st[-1] = i32(a << (b & 31))
                elif op == "ishr": st[-1] = a >> (b & 31)
                elif op == "iushr": st[-1] = i32((a & 0xFFFFFFFF) >> (b & 31))
                elif op == "iand": st[-1] = a & b
                elif op == "ior": st[-1] = a | b
                else: st[-1] = a ^ b
            elif op in ("dcmpl", "dcmpg", "fcmpl", "fcmpg"):
                b = st.pop(); a = st.pop()
                if a != a or b != b:
                    st.append(1 if op.endswith("g") else -1)
                else:
                    st.append(1 if a > b else (-1 if a < b else 0))
            elif op == "lcmp":
                b = int(st.pop()); a = int(st.pop())
                st.append(1 if a > b else (-1 if a < b else 0))
            elif c0 == "i" and op.startswith("if"):
                if op.startswith("if_icmp"):
                    b = st.pop(); a = st.pop(); cond = op[7:]
                elif op.startswith("if_acmp"):
                    b = st.pop(); a = st.pop()
                    taken = (a is b) if op.endswith("eq") else (a is not b)
                    if taken:
                        k = pcmap[ins.args[0]]
                    continue
                elif op in ("ifnull", "ifnonnull"):
                    a = st.pop()
                    if (a is None) == (op == "ifnull"):
                        k = pcmap[ins.args[0]]
                    continue
                else:
                    a = st.pop(); b = 0; cond = op[2:]
                taken = {"eq": a == b, "ne": a != b, "lt": a < b, "ge": a >= b, "gt": a > b, "le": a <= b}[cond]
                if taken:
                    k = pcmap[ins.args[0]]
            elif op in ("goto", "goto_w"):
                k = pcmap[ins.args[0]]
            elif op == "dup":
                st.append(st[-1])
            elif op == "dup2":
                if is_cat2(st[-1]):
                    st.append(st[-1])
                else:
                    st.extend(st[-2:])
            elif op == "dup_x1":
                v1 = st.pop(); v2 = st.pop()
                st.extend((v1, v2, v1))
            elif op == "dup_x2":
                v1 = st.pop(); v2 = st.pop()
                if is_cat2(v2):
                    st.extend((v1, v2, v1))
                else:
                    v3 = st.pop()
                    st.extend((v1, v3, v2, v1))
            elif op == "dup2_x1":
                v1 = st.pop()
                if is_cat2(v1):
                    v2 = st.pop()
                    st.extend((v1, v2, v1))
                else:
                    v2 = st.pop(); v3 = st.pop()
                    st.extend((v2, v1, v3, v2, v1))
            elif op == "dup2_x2":
                v1 = st.pop()
                if is_cat2(v1):
                    v2 = st.pop()
                    if is_cat2(v2):
                        st.extend((v1, v2, v1))
                    else:
                        v3 = st.pop()
                        st.extend((v1, v3, v2, v1))
                else:
                    v2 = st.pop(); v3 = st.pop()
                    if is_cat2(v3):
                        st.extend((v2, v1, v3, v2, v1))
                    else:
                        v4 = st.pop()
                        st.extend((v2, v1, v4, v3, v2, v1))
            elif op == "pop":
                st.pop()
            elif op == "pop2":
                if not is_cat2(st.pop()):
                    st.pop()
            elif op == "swap":
                st[-1], st[-2] = st[-2], st[-1]
            elif op in ("invokevirtual", "invokeinterface", "invokespecial", "invokestatic"):
                owner, name, desc = ins.ref
                params, ret = parse_desc(desc)
                n = len(params)
                a = st[len(st) - n:] if n else []
                if n:
                    del st[len(st) - n:]
                if op == "invokestatic":
                    r = self.call(owner, name, desc, a)
                else:
                    obj = st.pop()
                    if obj is None:
                        raise JavaThrow(f"NullPointerException invoking {owner}.{name}")
                    if op == "invokespecial":
                        nat, cf2, m2 = self._find(owner, name, desc)
                        if nat is not None:
                            r = nat(self, [obj] + a)
                        elif m2 is not None:
                            r = self.run(cf2, m2, [obj] + a)
                        else:
                            raise NotImplementedError(f"invokespecial {owner}.{name}{desc}")
                    else:
                        r = self.call_virtual(obj, owner, name, desc, a)
                if ret != "V":
                    if ret == "Z" and isinstance(r, bool):
                        r = int(r)
                    st.append(r)
            elif op in ("ireturn", "lreturn", "freturn", "dreturn", "areturn"):
                return st.pop()
            elif op == "return":
                return None
            elif op == "new":
                st.append(JObject(ins.ref[1]))
            elif op == "newarray":
                n = st.pop()
                t = ins.args[0]  # 4 bool 5 char 6 float 7 double 8 byte 9 short 10 int 11 long
                st.append([0.0] * n if t == 7 else [F32(0)] * n if t == 6 else [JLong(0)] * n if t == 11 else [0] * n)
            elif op == "anewarray":
                st.append([None] * st.pop())
            elif op == "multianewarray":
                dims = [st.pop() for _ in range(ins.args[1])][::-1]
                elem = ins.ref[1].lstrip("[")

                def make(d):
                    if d == len(dims) - 1:
                        z = 0.0 if elem == "D" else F32(0) if elem == "F" else JLong(0) if elem == "J" else 0 if elem in "IZBSC" else None
                        return [z] * dims[d]
                    return [make(d + 1) for _ in range(dims[d])]
                st.append(make(0))
            elif op == "checkcast":
                pass
            elif op == "instanceof":
                o = st.pop()
                st.append(int(o is not None and self._instanceof(o, ins.ref[1])))
            elif op == "athrow":
                raise JavaThrow(st.pop())
            elif op in ("monitorenter", "monitorexit"):
                st.pop()
            elif op in ("lookupswitch", "tableswitch"):
                v = st.pop()
                if op == "tableswitch":
                    dflt, lo, hi, tg = ins.args
                    k = pcmap[tg[v - lo]] if lo <= v <= hi else pcmap[dflt]
                else:
                    dflt, pairs = ins.args
                    k = pcmap[dict(pairs).get(v, dflt)]
            elif op == "nop":
                pass
            else:
                raise NotImplementedError(f"opcode {op} in {cf.name}.{m.name}")

    def _instanceof(self, o, cls: str) -> bool:
        c = o.cls if isinstance(o, JObject) else getattr(o, "jclass", None)
        while c is not None:
            if c == cls:
                return True
            if not self.jar.has(c):
                return False
            cf = self.jar.load(c)
            if cls in cf.interfaces:
                return True
            c = cf.super_name
        return False


class HostIterator:
    """java.util.Iterator over a Python sequence."""
    jclass = "java/util/Iterator"

    def __init__(self, seq):
        self.seq, self.i = seq, 0

    def has_next(self):
        return self.i < len(self.seq)

    def next(self):
        v = self.seq[self.i]
        self.i += 1
        return v
