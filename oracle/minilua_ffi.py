"""A LuaJIT-`ffi` look-alike for oracle/minilua.py, backed by ctypes. TEST INFRASTRUCTURE ONLY.

Purpose: EXECUTE the LuaJIT binding this project ships (`lua-multigrid-poisson_b200/lua/multigrid-poisson/cuda.lua`,
the drop-in class a reference maintainer would `require`) even though the image has no Lua runtime: minilua
interprets the Lua, this module gives it `require 'ffi'` -- cdef, load, new, gc, string, sizeof, cast-free pointer
passing -- so that `cl(size, real, cpuDepth):run()` (test/test.lua:53-56 of the reference) really drives
libmgpoisson.so through the C ABI from Lua source. Only the part of the FFI that binding uses is implemented:

  ffi.cdef(text)   enums (`enum { A = 0, ... };`), opaque struct typedefs, function prototypes with scalar and
                   pointer parameters (the MGPOISSON_CDEF block of include/mgpoisson.h)
  ffi.load(name)   dlopen; `lib.NAME` resolves enum constants and prototypes; calls convert arguments by the
                   declared parameter types (nil -> NULL, numbers -> int / double / size_t, cdata -> address)
  ffi.new(ct, ...) `T[n]`, `T[?]` (+ count), `T*[n]`; element access converts like LuaJIT (numbers in and out,
                   pointers as cdata)
  ffi.gc(p, fin)   registers a finaliser that runs when the cdata is collected or at FFI.close()
  ffi.string(p)    NUL-terminated C string -> Lua string
  ffi.sizeof(ct)
"""
import ctypes as C
import os
import re

import minilua as ml

_SCALARS = {"int": C.c_int, "unsigned int": C.c_uint, "unsigned": C.c_uint, "double": C.c_double, "float": C.c_float,
            "size_t": C.c_size_t, "uint64_t": C.c_uint64, "int64_t": C.c_int64, "uint32_t": C.c_uint32, "int32_t": C.c_int32,
            "char": C.c_char, "long": C.c_long, "unsigned long long": C.c_ulonglong}


class CPointer:
    """A pointer cdata (opaque)."""

    def __init__(self, addr, ffi=None):
        self.addr = int(addr or 0)
        self._fin = None
        self._ffi = ffi

    def run_finalizer(self):
        fin, self._fin = self._fin, None
        if fin is not None and self.addr:
            ml.lua_call(fin, [self])

    def __del__(self):
        try:
            self.run_finalizer()
        except Exception:
            pass

    def __eq__(self, other):
        if other is None:
            return self.addr == 0
        return isinstance(other, CPointer) and other.addr == self.addr

    def __hash__(self):
        return hash(self.addr)


class CArrayData:
    """`T[n]` or `T*[n]` cdata."""

    def __init__(self, elem, n, is_ptr):
        self.elem, self.n, self.is_ptr = elem, int(n), is_ptr
        self.buf = ((C.c_void_p if is_ptr else _SCALARS[elem]) * max(self.n, 1))()

    @property
    def addr(self):
        return C.addressof(self.buf)

    def lua_index(self, k):
        i = int(k)
        if not 0 <= i < self.n:
            raise ml.LuaError(f"cdata index {k!r} outside [0, {self.n})")
        v = self.buf[i]
        if self.is_ptr:
            return CPointer(v)
        return float(v) if not isinstance(v, bytes) else float(v[0])     # char: a number, like LuaJIT

    def lua_setindex(self, k, v):
        i = int(k)
        if not 0 <= i < self.n:
            raise ml.LuaError(f"cdata index {k!r} outside [0, {self.n})")
        if self.is_ptr:
            self.buf[i] = v.addr if v is not None else None
        elif self.elem in ("double", "float"):
            self.buf[i] = float(v)
        else:
            self.buf[i] = int(v)

    def numpy(self, dtype):
        import numpy as np
        return np.frombuffer(self.buf, dtype=dtype, count=self.n).copy()


class CTypedPointer(CPointer):
    """`T*` obtained by ffi.cast: indexable (from 0) memory somebody else owns."""

    def __init__(self, elem, addr):
        super().__init__(addr)
        self.elem = elem
        self.ct = _SCALARS[elem]

    def lua_index(self, k):
        v = self.ct.from_address(self.addr + int(k) * C.sizeof(self.ct)).value
        return float(v) if not isinstance(v, bytes) else float(v[0])

    def lua_setindex(self, k, v):
        self.ct.from_address(self.addr + int(k) * C.sizeof(self.ct)).value = float(v) if self.elem in ("double", "float") else int(v)


class _Proto:
    def __init__(self, ret, name, params):
        self.ret, self.name, self.params = ret, name, params


def _norm_type(t):
    t = re.sub(r"\bconst\b", " ", t)
    t = re.sub(r"\s+", " ", t).strip()
    t = re.sub(r"\s*\*\s*", "*", t)
    return t


class CLib:
    def __init__(self, ffi, cdll, path):
        self.ffi, self.cdll, self.path = ffi, cdll, path
        self._fns = {}

    def lua_index(self, k):
        if k in self.ffi.enums:
            return float(self.ffi.enums[k])
        if k in self.ffi.protos:
            if k not in self._fns:
                self._fns[k] = self._bind(self.ffi.protos[k])
            return self._fns[k]
        raise ml.LuaError(f"missing declaration for symbol '{k}'")

    def _bind(self, p):
        try:
            fn = getattr(self.cdll, p.name)
        except AttributeError:
            raise ml.LuaError(f"cannot resolve symbol '{p.name}' in {self.path}")
        ret = p.ret
        fn.restype = None if ret == "void" else (C.c_void_p if ret.endswith("*") else _SCALARS[ret])
        argtypes = [C.c_void_p if t.endswith("*") else _SCALARS[t] for t in p.params]
        fn.argtypes = argtypes

        def call(*args):
            if len(args) != len(p.params):
                raise ml.LuaError(f"wrong number of arguments for function call ({p.name}: got {len(args)}, want {len(p.params)})")
            conv = []
            for a, t in zip(args, p.params):
                if t.endswith("*"):
                    if a is None:
                        conv.append(None)
                    elif isinstance(a, (CPointer, CArrayData)):
                        conv.append(a.addr)
                    elif isinstance(a, str):
                        conv.append(C.cast(C.c_char_p(a.encode()), C.c_void_p))
                    elif hasattr(a, "a"):                      # run_reference.CArray (numpy backed)
                        conv.append(a.a.ctypes.data)
                    else:
                        raise ml.LuaError(f"cannot convert '{type(a).__name__}' to '{t}' ({p.name})")
                elif t in ("double", "float"):
                    conv.append(float(a))
                else:
                    if isinstance(a, float) and not a.is_integer():
                        a = int(a)     # LuaJIT truncates
                    conv.append(int(a))
            r = fn(*conv)
            if ret == "void":
                return None
            if ret.endswith("*"):
                return (CPointer(r),)
            return (float(r),)

        return call


class FFI:
    def __init__(self, search_dirs=()):
        self.enums, self.protos, self.structs = {}, {}, set()
        self.search_dirs = list(search_dirs)
        self.libs = []
        self._gc = []

    # ---- the module table for `require 'ffi'`
    def module(self):
        return ml.Interpreter.table_from({"cdef": self.cdef, "load": self.load, "new": self.new, "gc": self.gc,
                                          "string": self.string, "sizeof": self.sizeof, "cast": self.cast})

    def cdef(self, text):
        text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
        text = re.sub(r"//[^\n]*", " ", text)
        for m in re.finditer(r"enum\s*\w*\s*\{(.*?)\}\s*;", text, flags=re.S):
            nxt = 0
            for item in m.group(1).split(","):
                item = item.strip()
                if not item:
                    continue
                if "=" in item:
                    name, val = (x.strip() for x in item.split("=", 1))
                    nxt = int(val, 0)
                else:
                    name = item
                self.enums[name] = nxt
                nxt += 1
        text = re.sub(r"enum\s*\w*\s*\{.*?\}\s*;", " ", text, flags=re.S)
        for m in re.finditer(r"typedef\s+struct\s+(\w+)\s+(\w+)\s*;", text):
            self.structs.add(m.group(2))
        text = re.sub(r"typedef\s+struct\s+\w+\s+\w+\s*;", " ", text)
        for decl in text.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"^(.*?)(\w+)\s*\((.*)\)$", decl)
            if not m:
                raise ml.LuaError(f"ffi.cdef: cannot parse declaration '{decl}'")
            ret, name, plist = _norm_type(m.group(1)), m.group(2), m.group(3).strip()
            params = []
            if plist and plist != "void":
                for prm in plist.split(","):
                    prm = prm.strip()
                    mm = re.match(r"^(.*?)(\b\w+)?$", prm)
                    ty = prm
                    # drop the parameter name (last identifier) unless the whole thing is a type
                    mm = re.match(r"^(.*[\s\*])(\w+)$", prm)
                    if mm and _norm_type(mm.group(1)).rstrip("*") in set(_SCALARS) | self.structs | {"void", "char"}:
                        ty = mm.group(1)
                    params.append(self._check_type(_norm_type(ty), decl))
            self.protos[name] = _Proto(self._check_type(ret, decl), name, params)

    def _check_type(self, t, where):
        base = t.rstrip("*")
        if t.endswith("*"):
            if base in _SCALARS or base in self.structs or base in ("void", "char"):
                return t
        elif t in _SCALARS or t == "void":
            return t
        raise ml.LuaError(f"ffi.cdef: unsupported type '{t}' in '{where}'")

    def load(self, name):
        cands = [name] if (os.sep in name or name.endswith(".so")) else \
            [os.path.join(d, f"lib{name}.so") for d in self.search_dirs] + [f"lib{name}.so"]
        last = None
        for c in cands:
            try:
                lib = CLib(self, C.CDLL(c), c)
                self.libs.append(lib)
                return lib
            except OSError as e:
                last = e
        raise ml.LuaError(f"ffi.load: cannot load '{name}': {last}")

    def new(self, ct, *init):
        m = re.match(r"^\s*([\w ]+?)\s*(\*?)\s*\[\s*(\?|\d+)\s*\]\s*$", ct)
        if not m:
            raise ml.LuaError(f"ffi.new: unsupported ctype '{ct}'")
        elem, star, n = m.group(1), m.group(2), m.group(3)
        init = list(init)
        if n == "?":
            n = init.pop(0)
        if not star and elem not in _SCALARS:
            raise ml.LuaError(f"ffi.new: unknown element type '{elem}'")
        arr = CArrayData(elem, n, bool(star))
        for i, v in enumerate(init):
            arr.lua_setindex(i, v)
        return arr

    def cast(self, ct, v):
        """ffi.cast('T*', pointer): a typed view of memory the library owns"""
        m = re.match(r"^\s*(?:const\s+)?([\w ]+?)\s*\*\s*$", ct)
        if not m or m.group(1) not in _SCALARS:
            raise ml.LuaError(f"ffi.cast: unsupported ctype '{ct}'")
        addr = v.addr if isinstance(v, (CPointer, CArrayData)) else int(v or 0)
        return CTypedPointer(m.group(1), addr)

    def gc(self, cdata, fin):
        cdata._fin = fin
        self._gc.append(cdata)
        return cdata

    def string(self, p, n=None):
        addr = p.addr if isinstance(p, (CPointer, CArrayData)) else None
        if not addr:
            raise ml.LuaError("ffi.string: NULL pointer")
        return (C.string_at(addr) if n is None else C.string_at(addr, int(n))).decode(errors="replace")

    def sizeof(self, ct):
        ct = ct.strip()
        if ct.endswith("*"):
            return float(C.sizeof(C.c_void_p))
        if ct not in _SCALARS:
            raise ml.LuaError(f"ffi.sizeof: unknown type '{ct}'")
        return float(C.sizeof(_SCALARS[ct]))

    def close(self):
        """Runs the pending ffi.gc finalisers (what LuaJIT does when the state closes)."""
        for c in self._gc:
            c.run_finalizer()
        self._gc = []
