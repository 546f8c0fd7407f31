"""Minimal reader for R's XDR serialisation (.RData / .rda / .rds).

Used to load the bundled abn datasets (``ex1.dag.data``, ``var33``, ``FCV`` ...) and the
reference's golden caches (reference ``data/*.RData``, ``tests/testthat/testdata/*.Rdata``,
``inst/extdata/*.rds``) without an R installation.  It decodes the serialisation format
that R's ``save()`` / ``saveRDS()`` write (versions 2 and 3, XDR byte order, wrapped in
gzip / bzip2 / xz) into plain Python / numpy objects:

* atomic vectors  -> ``RVec`` (a numpy array subclass carrying ``.attrs``)
* generic vectors -> ``RList`` (a list subclass with ``.names`` / ``.attrs`` and name lookup)
* pairlists       -> ``RList`` as well (tags become names)
* factors stay INT vectors with ``attrs['levels']`` and ``attrs['class'] == ['factor']``

Language objects, closures, environments, byte-code and external pointers are consumed
structurally and replaced by ``None`` placeholders - nothing in the score-cache fixtures
needs them.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct

import numpy as np

NA_INT = -2147483648


class RVec(np.ndarray):
    """numpy array + R attributes (``names``, ``dim``, ``levels``, ``class`` ...)."""

    def __new__(cls, arr, attrs=None):
        obj = np.asarray(arr).view(cls)
        obj.attrs = attrs or {}
        return obj

    def __array_finalize__(self, obj):
        self.attrs = getattr(obj, "attrs", {})


class RList(list):
    """R generic vector / pairlist: positional list with optional names."""

    def __init__(self, items=(), names=None, attrs=None):
        super().__init__(items)
        self.attrs = attrs or {}
        if names is not None:
            self.attrs["names"] = list(names)

    @property
    def names(self):
        return self.attrs.get("names")

    def get(self, key, default=None):
        names = self.names or []
        return self[names.index(key)] if key in names else default

    def __getitem__(self, key):
        if isinstance(key, str):
            names = self.names or []
            if key not in names:
                raise KeyError(key)
            return list.__getitem__(self, names.index(key))
        return list.__getitem__(self, key)

    def keys(self):
        return list(self.names or [])


class _Placeholder:
    def __init__(self, kind):
        self.kind = kind

    def __repr__(self):
        return f"<R {self.kind}>"


def _decompress(raw: bytes) -> bytes:
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.p = 0
        self.refs = []

    # -- primitives -----------------------------------------------------------------
    def int(self) -> int:
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def length(self) -> int:
        n = self.int()
        if n == -1:
            hi, lo = self.int(), self.int()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def raw(self, n: int) -> bytes:
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def array(self, dtype: str, n: int) -> np.ndarray:
        a = np.frombuffer(self.b, dtype=dtype, count=n, offset=self.p)
        self.p += n * a.dtype.itemsize
        return a.astype(a.dtype.newbyteorder("="))

    # -- items ----------------------------------------------------------------------
    def attrs(self) -> dict:
        pl = self.item()
        out = {}
        if isinstance(pl, RList):
            for k, v in zip(pl.names or [], pl):
                out[k] = v
        return out

    def item(self):
        flags = self.int()
        t = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)

        if t == 254:  # NILVALUE
            return None
        if t in (253, 252, 251, 250, 242, 241):  # global env, unbound, missing, base ns, empty env, base env
            return _Placeholder(f"special{t}")
        if t == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if t in (249, 248, 247):  # NAMESPACESXP, PACKAGESXP, PERSISTSXP
            self.int()
            n = self.int()
            info = [self.item() for _ in range(n)]
            obj = _Placeholder(f"namespace{info}")
            self.refs.append(obj)
            return obj
        if t == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 4:  # ENVSXP
            self.int()  # locked
            obj = _Placeholder("environment")
            self.refs.append(obj)
            self.item()  # enclos
            self.item()  # frame
            self.item()  # hashtab
            self.item()  # attrib
            return obj
        if t in (2, 3, 5, 6, 17, 239, 240):  # pairlist-like
            items, names = [], []
            attrs = {}
            first = True
            while True:
                if t in (239, 240):  # ATTRLISTSXP / ATTRLANGSXP (byte-code consts)
                    has_attr, t = True, (2 if t == 239 else 6)
                if has_attr:
                    a = self.attrs()
                    if first:
                        attrs = a
                tag = self.item() if has_tag else None
                car = self.item()
                items.append(car)
                names.append(tag if isinstance(tag, str) else "")
                first = False
                # cdr: iterate instead of recursing
                flags = self.int()
                t2 = flags & 0xFF
                if t2 == 254:
                    break
                if t2 not in (2, 3, 5, 6, 17):
                    # dotted pair tail: re-parse as a normal item
                    self.p -= 4
                    items.append(self.item())
                    names.append("")
                    break
                t = t2
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
            return RList(items, names if any(names) else None, attrs)
        if t == 22:  # EXTPTRSXP
            obj = _Placeholder("externalptr")
            self.refs.append(obj)
            self.item()
            self.item()
            if has_attr:
                self.attrs()
            return obj
        if t == 23:  # WEAKREFSXP
            obj = _Placeholder("weakref")
            self.refs.append(obj)
            if has_attr:
                self.attrs()
            return obj
        if t in (7, 8):  # SPECIALSXP / BUILTINSXP
            n = self.int()
            return _Placeholder("builtin:" + self.raw(n).decode("latin1"))
        if t == 9:  # CHARSXP
            n = self.int()
            if n == -1:
                return None
            return self.raw(n).decode("utf-8", errors="replace")
        if t == 238:  # ALTREP
            info = self.item()
            state = self.item()
            attr = self.item()
            cls = info[0] if isinstance(info, RList) and len(info) else None
            out = self._altrep(cls, state)
            if isinstance(attr, RList) and isinstance(out, (RVec, RList)):
                for k, v in zip(attr.names or [], attr):
                    out.attrs[k] = v
            return out
        if t == 25:  # S4SXP
            return RList([], None, self.attrs() if has_attr else {})
        if t == 21:  # BCODESXP
            raise NotImplementedError("byte-code objects are not supported")

        # vectors ---------------------------------------------------------------------
        if t in (10, 13):
            n = self.length()
            out = RVec(self.array(">i4", n))
        elif t == 14:
            n = self.length()
            out = RVec(self.array(">f8", n))
        elif t == 15:
            n = self.length()
            out = RVec(self.array(">f8", 2 * n).view(np.complex128))
        elif t == 16:
            n = self.length()
            out = RList([self.item() for _ in range(n)])
            out.attrs["_str"] = True
        elif t in (19, 20):
            n = self.length()
            out = RList([self.item() for _ in range(n)])
        elif t == 24:
            n = self.length()
            out = RVec(np.frombuffer(self.raw(n), dtype=np.uint8))
        else:
            raise NotImplementedError(f"SEXP type {t} at offset {self.p}")
        if t == 10:
            out.attrs["_logical"] = True
        if has_attr:
            out.attrs.update(self.attrs())
        return out

    def _altrep(self, cls, state):
        if cls == "compact_intseq":
            n, start, step = (int(x) for x in state[:3])
            return RVec(np.arange(start, start + n * step, step, dtype=np.int32))
        if cls == "compact_realseq":
            n, start, step = state[:3]
            return RVec(start + step * np.arange(int(n), dtype=np.float64))
        if cls in ("wrap_real", "wrap_integer", "wrap_logical", "wrap_string", "wrap_list"):
            return state[0]
        if cls == "deferred_string":
            src = state[0]
            return RList([None if (isinstance(x, (int, np.integer)) and x == NA_INT) else _fmt(x) for x in src])
        raise NotImplementedError(f"ALTREP class {cls}")


def _fmt(x):
    if isinstance(x, (float, np.floating)) and float(x).is_integer():
        return str(int(x))
    return str(x)


def _header(rd: _Reader):
    fmt = rd.raw(2)
    if fmt != b"X\n":
        raise ValueError("only XDR serialisation is supported")
    version = rd.int()
    rd.int()
    rd.int()
    if version == 3:
        n = rd.int()
        rd.raw(n)
    elif version != 2:
        raise ValueError(f"unsupported serialisation version {version}")


def strings(x) -> list:
    """STRSXP -> list[str|None]."""
    return list(x) if x is not None else []


def read_rds(path: str):
    """Read a ``saveRDS`` file and return the single object it holds."""
    rd = _Reader(_decompress(open(path, "rb").read()))
    _header(rd)
    return rd.item()


def read_rdata(path: str) -> dict:
    """Read a ``save`` / ``save.image`` file and return ``{name: object}``."""
    buf = _decompress(open(path, "rb").read())
    if buf[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError(f"{path}: not an RDX2/RDX3 file")
    rd = _Reader(buf)
    rd.p = 5
    _header(rd)
    top = rd.item()
    return {k: v for k, v in zip(top.names, top)}


def is_factor(col) -> bool:
    cls = getattr(col, "attrs", {}).get("class")
    return cls is not None and "factor" in list(cls)


def data_frame(obj: RList):
    """R data.frame -> (names, columns).

    Each column is either a float64 / int32 numpy array or, for factors, a tuple
    ``(codes int32 1-based, levels list[str])``.
    """
    names = list(obj.names)
    cols = []
    for c in obj:
        if is_factor(c):
            cols.append((np.asarray(c, dtype=np.int32), list(c.attrs["levels"])))
        elif isinstance(c, RList):
            cols.append(list(c))
        else:
            cols.append(np.asarray(c))
    return names, cols


def matrix(obj: RVec) -> np.ndarray:
    """R matrix (column-major with ``dim``) -> 2-D numpy array."""
    dim = [int(x) for x in obj.attrs["dim"]]
    return np.asarray(obj).reshape(dim, order="F")
