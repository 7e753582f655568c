"""f90_runtime.py -- TEST INFRASTRUCTURE (never imported by the product path).

Run-time support for the Python text that `oracle/f90_translate.py` produces from the reference's own Fortran source
files.  It fixes the semantics the translated statements rely on:

* scalars: Fortran `integer` = Python int, `real` = numpy.float32, `double precision` = numpy.float64.  NumPy 2's
  promotion rules (NEP 50) then give Fortran's mixed-mode arithmetic: int*real -> real, real*double -> double with the
  single-precision value WIDENED (so the literal 0.0000001 in a double expression is float(np.float32(1e-7)), as a Fortran
  compiler evaluates it), int/int truncates (`f_div`).
* exp / log / cos go through the C library (`math`), the same libm a gfortran build of the reference calls.
* arrays: `FArr`, 1-based (or declared lower bounds), column-major index order, sections are views, whole-array
  arithmetic elementwise, `sum` left to right in element order (what gfortran inlines without -ffast-math).
* arguments: arrays and derived types by reference (Python objects); scalar dummies by copy-in/copy-out, which equals
  by-reference whenever no two arguments alias -- true of every call on the path.
"""
import copy
import math
import os

import numpy as np

F32 = np.float32
F64 = np.float64
D0 = F64(0.0)


class FStop(Exception):
    """A Fortran STOP statement was executed."""


class FUnsupported(Exception):
    """A statement or procedure outside the translated subset was reached at run time."""


class IOLog:
    """print / write statements are not formatted: the first string literal of the I/O list names the event and the
    values of the other list items are kept (the reference's warnings -- ' Storebal SD wrong KINEMATIC ', ' Problem with EX '
    -- and the numbers it writes to the .res / .flow units)."""

    def __init__(self, cap=2000000):
        self.events = []
        self.counts = {}
        self.cap = cap

    def __call__(self, where, text, values):
        self.counts[text] = self.counts.get(text, 0) + 1
        if len(self.events) < self.cap:
            vals = [np.array(v.a, copy=True) if isinstance(v, FArr) else v for v in values]
            self.events.append((where, text, vals))

    def clear(self):
        self.events.clear()
        self.counts.clear()


IO = IOLog()


def f_io(where, text, values=()):
    IO(where, text, values)


def f_unsupported(where, text):
    raise FUnsupported(f"{where}: {text}")


def f_stop(where, text):
    raise FStop(f"{where}: STOP {text}")


# ------------------------------------------------------------------------------------------------ arrays
class FArr:
    __slots__ = ("a", "lb", "isint", "n0", "l0")

    def __init__(self, a, lb=None):
        self.a = a
        self.lb = tuple(lb) if lb is not None else (1,) * a.ndim
        self.isint = a.dtype.kind in "iu"
        self.n0 = a.shape[0] if a.ndim else 0
        self.l0 = self.lb[0] if a.ndim else 1

    # -- indexing ---------------------------------------------------------------------------------
    def _key(self, k):
        if not isinstance(k, tuple):
            k = (k,)
        if len(k) != self.a.ndim:
            raise IndexError(f"rank mismatch: {len(k)} subscripts for rank {self.a.ndim}")
        out = []
        sec = False
        for kk, l, n in zip(k, self.lb, self.a.shape):
            if isinstance(kk, slice):
                lo = l if kk.start is None else int(kk.start)
                hi = l + n - 1 if kk.stop is None else int(kk.stop)
                st = 1 if kk.step is None else int(kk.step)
                if st != 1:
                    raise FUnsupported("strided section")
                if hi < lo:
                    out.append(slice(0, 0))
                else:
                    if lo < l or hi > l + n - 1:
                        raise IndexError(f"section {lo}:{hi} outside {l}:{l + n - 1}")
                    out.append(slice(lo - l, hi - l + 1))
                sec = True
            else:
                i = int(kk) - l
                if i < 0 or i >= n:
                    raise IndexError(f"subscript {kk} outside {l}:{l + n - 1}")
                out.append(i)
        return tuple(out), sec

    def __getitem__(self, k):
        if type(k) is int and self.a.ndim == 1:
            i = k - self.l0
            if i < 0 or i >= self.n0:
                raise IndexError(f"subscript {k} outside {self.l0}:{self.l0 + self.n0 - 1}")
            v = self.a[i]
            return int(v) if self.isint else v
        key, sec = self._key(k)
        v = self.a[key]
        if sec:
            return FArr(v)
        return int(v) if self.isint else v

    def __setitem__(self, k, v):
        if type(k) is int and self.a.ndim == 1:
            i = k - self.l0
            if i < 0 or i >= self.n0:
                raise IndexError(f"subscript {k} outside {self.l0}:{self.l0 + self.n0 - 1}")
            self.a[i] = v
            return
        key, sec = self._key(k)
        if isinstance(v, FArr):
            v = v.a
        self.a[key] = v

    # -- whole-array arithmetic (elementwise IEEE operations) ---------------------------------------
    @staticmethod
    def _u(x):
        return x.a if isinstance(x, FArr) else x

    def __add__(self, o): return FArr(self.a + FArr._u(o))
    def __radd__(self, o): return FArr(FArr._u(o) + self.a)
    def __sub__(self, o): return FArr(self.a - FArr._u(o))
    def __rsub__(self, o): return FArr(FArr._u(o) - self.a)
    def __mul__(self, o): return FArr(self.a * FArr._u(o))
    def __rmul__(self, o): return FArr(FArr._u(o) * self.a)
    def __neg__(self): return FArr(-self.a)

    def __truediv__(self, o):
        u = FArr._u(o)
        if self.isint and (isinstance(u, int) or getattr(u, "dtype", np.dtype("f8")).kind in "iu"):
            raise FUnsupported("integer array division")
        return FArr(self.a / u)

    def __rtruediv__(self, o):
        if self.isint and isinstance(o, int):
            raise FUnsupported("integer array division")
        return FArr(FArr._u(o) / self.a)

    def tolist(self):
        return self.a.tolist()

    def __len__(self):
        return self.a.size

    def __repr__(self):
        return f"FArr(lb={self.lb}, {self.a!r})"


_DT = {"i": np.int64, "r": np.float32, "d": np.float64, "l": np.bool_, "c": object, "t": object}


def f_alloc(base, bounds, factory=None):
    """bounds: list of (lo, hi).  Derived-type arrays are filled with fresh instances."""
    lb = [int(b[0]) for b in bounds]
    shape = [max(int(b[1]) - int(b[0]) + 1, 0) for b in bounds]
    a = np.zeros(shape, dtype=_DT[base], order="F")
    if base == "t":
        flat = a.reshape(-1, order="F") if a.size else a
        for i in range(a.size):
            flat[i] = factory()
        a = flat.reshape(shape, order="F") if a.size else a
    elif base == "c":
        a[...] = ""
    return FArr(a, lb)


def f_aassign(dst, src):
    """whole-array assignment `a = expr` (scalar broadcast or conforming array; allocates an unallocated allocatable)."""
    if dst is None:
        if isinstance(src, FArr):
            return FArr(np.array(src.a, order="F", copy=True))
        raise FUnsupported("assignment of a scalar to an unallocated array")
    if isinstance(src, FArr):
        if src.a.shape != dst.a.shape:
            raise IndexError(f"shape mismatch {src.a.shape} -> {dst.a.shape}")
        dst.a[...] = src.a
    else:
        dst.a[...] = src
    return dst


def f_gather(arr, comp):
    """`x%c` where x is an array of derived type: the array of that component (a copy, read-only use)."""
    flat = arr.a.reshape(-1, order="F")
    vals = [getattr(e, comp) for e in flat]
    kind = "d"
    if vals and isinstance(vals[0], (int, np.integer)) and not isinstance(vals[0], bool):
        kind = "i"
    out = np.array(vals, dtype=_DT[kind]).reshape(arr.a.shape, order="F")
    return FArr(out, arr.lb)


def f_scatter(arr, comp, value, conv):
    flat = arr.a.reshape(-1, order="F")
    if isinstance(value, FArr):
        src = value.a.reshape(-1, order="F")
        for e, v in zip(flat, src):
            setattr(e, comp, conv(v))
    else:
        for e in flat:
            setattr(e, comp, conv(value))


def f_slice(lo, hi, step=None):
    return slice(lo, hi, step)


def f_range(a, b, c=1):
    a, b, c = int(a), int(b), int(c)
    return range(a, b + (1 if c > 0 else -1), c)


# ------------------------------------------------------------------------------------------------ scalars
def cv_d(x):
    return F64(x)


def cv_r(x):
    return F32(x)


def cv_i(x):
    if isinstance(x, (float, np.floating)):
        return int(x)               # truncation toward zero, as Fortran's implicit conversion
    return int(x)


def cv_l(x):
    return bool(x)


def cv_c(x):
    return x


def cv_t(x):
    return copy.deepcopy(x)


def _isint(x):
    return isinstance(x, (int, np.integer)) and not isinstance(x, (bool, np.bool_))


def f_div(a, b):
    if isinstance(a, FArr) or isinstance(b, FArr):
        return a / b if isinstance(a, FArr) else b.__rtruediv__(a)
    if _isint(a) and _isint(b):
        a, b = int(a), int(b)
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    if _isint(a):
        a = type(b)(a)
    if _isint(b):
        b = type(a)(b)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        return a / b


def f_pow(a, b):
    if _isint(b):
        if _isint(a):
            return int(a) ** int(b)
        # libgcc __powidf2 (what gfortran emits for real**integer)
        n = abs(int(b))
        x = a
        y = x if n % 2 else type(a)(1)
        n >>= 1
        while n:
            x = x * x
            if n % 2:
                y = y * x
            n >>= 1
        return type(a)(1) / y if b < 0 else y
    t = F64 if isinstance(a, F64) or isinstance(b, F64) else F32
    return t(math.pow(float(a), float(b)))


def _common(args):
    if any(isinstance(x, F64) for x in args):
        return F64
    if any(isinstance(x, F32) for x in args):
        return F32
    return int


def i_min(*a):
    t = _common(a)
    return min(t(x) for x in a)


def i_max(*a):
    t = _common(a)
    return max(t(x) for x in a)


def _fl(fn):
    def g(x):
        t = F32 if isinstance(x, F32) else F64
        try:
            return t(fn(float(x)))
        except OverflowError:
            return t(math.inf)
        except ValueError:
            return t(math.nan)
    return g


i_exp = _fl(math.exp)
i_cos = _fl(math.cos)
i_sin = _fl(math.sin)
i_tan = _fl(math.tan)
i_atan = _fl(math.atan)
i_sqrt = _fl(math.sqrt)


def i_log(x):
    t = F32 if isinstance(x, F32) else F64
    xf = float(x)
    if xf > 0.0:
        return t(math.log(xf))
    if xf == 0.0:
        return t(-math.inf)
    return t(math.nan)


def i_abs(x):
    if isinstance(x, FArr):
        return FArr(np.abs(x.a))
    return abs(x)


def i_dble(x):
    return F64(x)


def i_real(x, kind=None):
    return F64(x) if kind == 8 else F32(x)


def i_int(x, kind=None):
    if isinstance(x, FArr):
        return FArr(np.trunc(x.a).astype(np.int64), x.lb)
    return int(x)


def i_nint(x):
    xf = float(x)
    r = math.floor(abs(xf))
    if abs(xf) - r >= 0.5:
        r += 1
    return int(r) if xf >= 0 else -int(r)


def i_floor(x):
    return int(math.floor(float(x)))


def i_ceiling(x):
    return int(math.ceil(float(x)))


def i_mod(a, b):
    if _isint(a) and _isint(b):
        return int(math.fmod(int(a), int(b)))
    t = _common((a, b))
    return t(math.fmod(float(a), float(b)))


def i_sign(a, b):
    return abs(a) if b >= 0 else -abs(a)


def i_size(x, dim=None):
    if dim is None:
        return int(x.a.size)
    return int(x.a.shape[int(dim) - 1])


def i_allocated(x):
    return x is not None


def i_sum(x, mask=None):
    """SUM(array [, mask]): left to right in array element order"""
    flat = x.a.reshape(-1, order="F")
    mk = None if mask is None else mask.a.reshape(-1, order="F")
    if x.isint:
        return int(flat.sum() if mk is None else flat[mk].sum())
    s = flat.dtype.type(0)
    if mk is None:
        for v in flat:
            s = s + v
    else:
        for v, keep in zip(flat, mk):
            if keep:
                s = s + v
    return s


def i_minval(x):
    flat = x.a.reshape(-1, order="F")
    v = flat.min()
    return int(v) if x.isint else v


def i_maxval(x):
    flat = x.a.reshape(-1, order="F")
    v = flat.max()
    return int(v) if x.isint else v


def i_cshift(x, shift, dim=1):
    return FArr(np.roll(x.a, -int(shift), axis=int(dim) - 1), x.lb)


def i_trim(x):
    return str(x).rstrip()


def i_len_trim(x):
    return len(str(x).rstrip())


def i_adjustl(x):
    return str(x).lstrip()


def i_huge(x):
    if _isint(x):
        return 2147483647
    return F64(np.finfo(np.float64).max) if isinstance(x, F64) else F32(np.finfo(np.float32).max)


def i_isnan(x):
    return bool(np.isnan(x))


def i_move_alloc(src, dst):
    raise FUnsupported("move_alloc is handled by the translator")


class Namespace:
    """module variables and named constants of the translated modules"""


def from_numpy(a, kind=None):
    """helper for drivers: numpy array -> FArr (column-major copy)"""
    a = np.asarray(a)
    if kind is None:
        kind = "i" if a.dtype.kind in "iu" else "d"
    return FArr(np.array(a, dtype=_DT[kind], order="F", copy=True))


# ------------------------------------------------------------------------------------------------ array predicates
def _cmp(op):
    def f(self, o):
        return FArr(op(self.a, FArr._u(o)))
    return f


FArr.__eq__ = _cmp(np.equal)
FArr.__ne__ = _cmp(np.not_equal)
FArr.__lt__ = _cmp(np.less)
FArr.__le__ = _cmp(np.less_equal)
FArr.__gt__ = _cmp(np.greater)
FArr.__ge__ = _cmp(np.greater_equal)
FArr.__hash__ = None


def i_count(mask):
    return int(np.count_nonzero(mask.a))


def f_case(sel, *vals):
    """SELECT CASE match; character selectors compare without trailing blanks"""
    if isinstance(sel, str):
        return any(sel.rstrip() == str(v).rstrip() for v in vals)
    return any(sel == v for v in vals)


# ------------------------------------------------------------------------------------------------ sequential formatted input
class Unit:
    def __init__(self, path):
        with open(path, "r", errors="replace") as f:
            self.lines = f.read().split("\n")
        if self.lines and self.lines[-1] == "":
            self.lines.pop()
        self.pos = 0
        self.path = path


UNITS = {}
OUTPUT_FILES = []      # (unit, name) of every file opened for writing


class FEndOfFile(Exception):
    pass


def open_unit(unit, path):
    UNITS[int(unit)] = Unit(path)


def f_open(where, unit, iostat=False, **kw):
    """-> the IOSTAT value.  Output files are not produced (their WRITEs are logged instead)."""
    if kw.get("action", "").strip().lower() == "write" or kw.get("status", "").strip().lower() in ("unknown", "replace", "new"):
        OUTPUT_FILES.append((int(unit), str(kw.get("file", "")).strip()))
        return 0
    path = str(kw["file"]).strip()
    if not os.path.isfile(path):
        if iostat:
            return 2
        raise FStop(f"{where}: cannot open {path!r}")
    open_unit(unit, path)
    return 0


def f_close(unit):
    UNITS.pop(int(unit), None)


def f_rewind(unit):
    UNITS[int(unit)].pos = 0


def _tokens(line):
    """list-directed values of one record: blank / comma / tab separated, quoted strings kept whole, '/' ends the record"""
    out = []
    i, n = 0, len(line)
    slash = False
    while i < n:
        ch = line[i]
        if ch in " \t\r":
            i += 1
        elif ch == ",":
            i += 1
        elif ch == "/":
            slash = True
            break
        elif ch in "'\"":
            j = i + 1
            buf = []
            while j < n:
                if line[j] == ch:
                    if j + 1 < n and line[j + 1] == ch:
                        buf.append(ch)
                        j += 2
                        continue
                    break
                buf.append(line[j])
                j += 1
            out.append("".join(buf))
            i = j + 1
        else:
            j = i
            while j < n and line[j] not in " \t\r,":
                j += 1
            out.append(line[i:j])
            i = j
    return out, slash


def _convert(tok, base, where):
    if base == "c":
        return tok
    t = tok.strip().lower()
    try:
        if base == "i":
            return int(t)
        v = float(t.replace("d", "e"))
        return F64(v) if base == "d" else F32(v)
    except ValueError:
        raise FStop(f"{where}: bad {base} value {tok!r} in list-directed input") from None


def _formatted_fields(line, fmt, need, where):
    """fields of one record under a format of A / Aw / Iw / nX descriptors (dyna_project.f90:330-331: a13,1X,a700 and
    a13,1X,i4); a record shorter than the format is padded with blanks, as Fortran does"""
    body = fmt.strip().lower()
    if not (body.startswith("(") and body.endswith(")")):
        raise FUnsupported(f"{where}: format {fmt}")
    out = []
    col = 0
    for d in [x.strip() for x in body[1:-1].split(",")]:
        if d == "a":
            out.append(line[col:])
            col = len(line)
        elif d[0] in "ai" and d[1:].isdigit():
            w = int(d[1:])
            out.append(line[col:col + w].ljust(w) if d[0] == "a" else (line[col:col + w].strip() or "0"))
            col += w
        elif d.endswith("x") and d[:-1].isdigit():
            col += int(d[:-1])
        else:
            raise FUnsupported(f"{where}: edit descriptor {d!r}")
    if len(out) < need:
        raise FUnsupported(f"{where}: format {fmt} holds fewer items than the list")
    return out


def f_read(where, unit, fmt, spec, iostat=False):
    """READ(unit, fmt) of the items described by spec = [(base, count or None)]; returns (io, [values]) where an array
    item's value is a list.  fmt: '*' (list directed) or '(a)' (one whole record into a character item).  Every READ
    starts a new record and abandons the rest of the last record it touched (Fortran sequential formatted input)."""
    u = UNITS.get(int(unit))
    if u is None:
        raise FStop(f"{where}: unit {unit} is not connected")
    need = sum(1 if c is None else int(c) for _, c in spec)
    flat = []
    try:
        if fmt != "*":
            if u.pos >= len(u.lines):
                raise FEndOfFile
            flat = _formatted_fields(u.lines[u.pos], fmt, need, where)
            u.pos += 1
        elif need == 0:
            if u.pos >= len(u.lines):
                raise FEndOfFile
            u.pos += 1
        else:
            while len(flat) < need:
                if u.pos >= len(u.lines):
                    raise FEndOfFile
                toks, slash = _tokens(u.lines[u.pos])
                u.pos += 1
                flat.extend(toks)
                if slash:
                    break
            if len(flat) < need:
                raise FUnsupported(f"{where}: '/' terminated input")
    except FEndOfFile:
        if iostat:
            return -1, None
        raise FStop(f"{where}: end of file on unit {unit} ({u.path})") from None
    vals = []
    k = 0
    for base, c in spec:
        if c is None:
            vals.append(_convert(flat[k], base, where))
            k += 1
        else:
            vals.append([_convert(t, base, where) for t in flat[k:k + int(c)]])
            k += int(c)
    return 0, vals


def f_store(dst, values, base):
    """whole-array / section target of a READ"""
    a = np.array(values, dtype=_DT[base] if base != "c" else object)
    if isinstance(dst, FArr):
        dst.a[...] = a.reshape(dst.a.shape, order="F")
        return dst
    raise FUnsupported("READ into an unallocated array")


def f_backspace(unit):
    u = UNITS[int(unit)]
    u.pos = max(u.pos - 1, 0)


# ------------------------------------------------------------------------------------------------ characters, command line
ARGV = []          # argument 0 is the program name, as GET_COMMAND_ARGUMENT numbers them


def f_get_command_argument(number):
    n = int(number)
    return ARGV[n] if 0 <= n < len(ARGV) else ""


def f_substr(s, lo, hi):
    lo = 1 if lo is None else int(lo)
    hi = len(s) if hi is None else int(hi)
    return s[lo - 1:hi].ljust(max(hi - lo + 1, 0)) if hi >= lo else ""


def f_setsub(s, lo, hi, value):
    lo = 1 if lo is None else int(lo)
    hi = len(s) if hi is None else int(hi)
    s = s.ljust(hi)
    w = hi - lo + 1
    return s[:lo - 1] + str(value)[:w].ljust(w) + s[hi:]


def i_len(s):
    return len(s)


def i_index(s, sub):
    return s.find(sub) + 1


def i_iachar(c):
    return ord(c[0]) if c else 32


i_ichar = i_iachar


def i_achar(k):
    return chr(int(k))


i_char = i_achar


def f_fmt(where, fmt, values):
    """internal WRITE under a format of character-literal, A / Aw, I0 / Iw / Iw.m and nX descriptors (no reversion)"""
    body = str(fmt).strip()
    if not (body.startswith("(") and body.endswith(")")):
        raise FUnsupported(f"{where}: format {fmt!r}")
    body = body[1:-1]
    parts, cur, q = [], [], None
    for ch in body:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == ",":
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    out = []
    vals = list(values)
    for d in parts:
        if not d:
            continue
        if d[0] in "'\"":
            out.append(d[1:-1].replace(d[0] * 2, d[0]))
            continue
        dl = d.lower()
        if dl.endswith("x") and dl[:-1].isdigit():
            out.append(" " * int(dl[:-1]))
        elif dl == "a":
            out.append(str(vals.pop(0)))
        elif dl[0] == "a" and dl[1:].isdigit():
            w = int(dl[1:])
            v = str(vals.pop(0))
            out.append(v[:w] if len(v) >= w else v.rjust(w))
        elif dl[0] == "i":
            w, _, m = dl[1:].partition(".")
            v = int(vals.pop(0))
            txt = str(abs(v)).rjust(int(m), "0") if m else str(abs(v))
            if v < 0:
                txt = "-" + txt
            w = int(w)
            if w and len(txt) > w:
                txt = "*" * w
            out.append(txt.rjust(w) if w else txt)
        else:
            raise FUnsupported(f"{where}: edit descriptor {d!r}")
    if vals:
        raise FUnsupported(f"{where}: format {fmt!r} has fewer descriptors than items")
    return "".join(out)


def i_iand(a, b):
    return int(a) & int(b)


def i_ior(a, b):
    return int(a) | int(b)


def i_ieor(a, b):
    return int(a) ^ int(b)


def i_ishft(a, n):
    return int(a) << int(n) if n >= 0 else int(a) >> -int(n)


def i_btest(a, pos):
    return bool((int(a) >> int(pos)) & 1)


# ------------------------------------------------------------------------------------------------ iso_c_binding
C_SIZEOF = {}      # class name of a bind(C) derived type -> sizeof of the C struct (filled by the driver from ctypes)


def i_c_loc(x):
    return x       # the "address" of an array is the array object itself: the callee sees the same storage


def i_c_sizeof(x):
    return C_SIZEOF[type(x).__name__]


def i_c_associated(x):
    return x is not None and not (isinstance(x, int) and x == 0)
