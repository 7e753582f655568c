"""f90_translate.py -- TEST INFRASTRUCTURE (never imported by the product path).

No Fortran compiler exists in the build image, so the reference cannot be RUN here.  This module is the next best pin:
it reads the reference's own source files where they lie (`/root/reference/RRMODEL/*.f90`, `DTA/*.f90`) and translates
the Fortran subset they use -- statement by statement, mechanically, with no knowledge of hydrology -- into Python text
that `oracle/f90_runtime.py` executes with Fortran's arithmetic semantics (typed literals, mixed-mode promotion, integer
division, libm exp/log, 1-based column-major arrays, copy-in/copy-out scalars).  `tests/golden/make_f90_vectors.py` runs
the translated `param_setup` / `init_satzone` / `Topmod` / `route_process_run_step` ... on small catchments and commits
the outputs; the C++ oracle must reproduce them bit for bit (`tests/test_f90_pin_cpu.py`).  What this pins: the oracle's
READING of the source (operation order, literal kinds, branch structure, quirks such as `.lt. - 0000000001`).  What it
cannot pin: a particular compiler's code generation (FMA contraction, x87, -ffast-math) -- DESIGN.md section 2.

Subset: modules, derived types (allocatable components, nesting), parameters, module variables, generic interfaces
(resolved on the first argument's type and rank), subroutines / functions (result clause), block and one-line IF,
DO (counted, labelled, WHILE), EXIT / CYCLE / RETURN / STOP, ALLOCATE / DEALLOCATE, array sections and whole-array
assignment, component access through arrays of derived type (`x%c = 0`), the intrinsics in f90_runtime.  PRINT / WRITE
are logged, not formatted; any other I/O statement raises if reached.  Nothing here is copied from the reference: the
output text exists only in memory (or wherever a caller chooses to dump it for debugging, outside the repository).
"""
import keyword
import re

# ================================================================================================== source lines

def logical_lines(text):
    """comment-stripped, continuation-joined, ';'-split, lower-cased (outside strings) statements -> [(line, text)]"""
    out = []
    cur = ""
    cur_ln = 0
    for ln, raw in enumerate(text.split("\n"), 1):
        s = []
        q = None
        for ch in raw:
            if q:
                s.append(ch)
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
                s.append(ch)
            elif ch == "!":
                break
            else:
                s.append(ch.lower())
        s = "".join(s).strip()
        if not s:
            continue
        if cur:
            if s.startswith("&"):
                s = s[1:].lstrip()
            cur += " " + s
        else:
            cur = s
            cur_ln = ln
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        part = []
        q = None
        for ch in cur:
            if q:
                part.append(ch)
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
                part.append(ch)
            elif ch == ";":
                p = "".join(part).strip()
                if p:
                    out.append((cur_ln, p))
                part = []
            else:
                part.append(ch)
        p = "".join(part).strip()
        if p:
            out.append((cur_ln, p))
        cur = ""
    return out


# ================================================================================================== tokens

_DOT = re.compile(r"\.(eq|ne|lt|le|gt|ge|and|or|not|eqv|neqv|true|false)\.")
_IDENT = re.compile(r"[a-z_][a-z0-9_]*")
_OPS = ["**", "//", "==", "/=", "<=", ">=", "=>", "::", "(/", "/)", "<", ">", "+", "-", "*", "/", "(", ")", ",", ":", "%", "=", "[", "]"]


class Tok:
    __slots__ = ("k", "v", "pos", "end")

    def __init__(self, k, v, pos, end):
        self.k, self.v, self.pos, self.end = k, v, pos, end

    def __repr__(self):
        return f"{self.k}:{self.v}"


def tokenize(s):
    toks = []
    i = 0
    n = len(s)
    while i < n:
        ch = s[i]
        if ch.isspace():
            i += 1
            continue
        if ch in "'\"":
            j = i + 1
            buf = []
            while j < n:
                if s[j] == ch:
                    if j + 1 < n and s[j + 1] == ch:
                        buf.append(ch)
                        j += 2
                        continue
                    break
                buf.append(s[j])
                j += 1
            toks.append(Tok("str", "".join(buf), i, j + 1))
            i = j + 1
            continue
        if ch == ".":
            m = _DOT.match(s, i)
            if m:
                toks.append(Tok("dot", m.group(1), i, m.end()))
                i = m.end()
                continue
        if ch.isdigit() or (ch == "." and i + 1 < n and s[i + 1].isdigit()):
            j = i
            isreal = False
            dbl = False
            while j < n and s[j].isdigit():
                j += 1
            if j < n and s[j] == "." and not _DOT.match(s, j):
                isreal = True
                j += 1
                while j < n and s[j].isdigit():
                    j += 1
            if j < n and s[j] in "ed":
                m = re.compile(r"[ed][+-]?\d+").match(s, j)
                if m:
                    isreal = True
                    dbl = s[j] == "d"
                    j = m.end()
            txt = s[i:j]
            if j < n and s[j] == "_":                       # kind suffix
                m = _IDENT.match(s, j + 1) or re.compile(r"\d+").match(s, j + 1)
                if m:
                    dbl = dbl or m.group(0) in ("8", "dp")
                    j = m.end()
            toks.append(Tok("num", (txt, "d" if dbl else ("r" if isreal else "i")), i, j))
            i = j
            continue
        m = _IDENT.match(s, i)
        if m:
            toks.append(Tok("id", m.group(0), i, m.end()))
            i = m.end()
            continue
        for op in _OPS:
            if s.startswith(op, i):
                # "(/" is an array constructor only when not followed by "=" ... keep "(" "/=" apart
                if op == "(/" and s.startswith("(/=", i):
                    continue
                toks.append(Tok("op", op, i, i + len(op)))
                i += len(op)
                break
        else:
            raise SyntaxError(f"cannot tokenize {s[i:i + 20]!r} in {s!r}")
    return toks


# ================================================================================================== expressions

class Parser:
    def __init__(self, toks, i=0):
        self.t = toks
        self.i = i

    def peek(self, k=0):
        j = self.i + k
        return self.t[j] if j < len(self.t) else None

    def at(self, kind, val=None):
        t = self.peek()
        return t is not None and t.k == kind and (val is None or t.v == val)

    def eat(self, kind, val=None):
        t = self.peek()
        if t is None or t.k != kind or (val is not None and t.v != val):
            raise SyntaxError(f"expected {kind} {val}, got {t} in {self.t}")
        self.i += 1
        return t

    def done(self):
        return self.i >= len(self.t)

    # precedence climbing, lowest first
    def expr(self):
        a = self.p_or()
        while self.at("dot", "eqv") or self.at("dot", "neqv"):
            op = self.eat("dot").v
            b = self.p_or()
            a = ("bin", op, a, b)
        return a

    def p_or(self):
        a = self.p_and()
        while self.at("dot", "or"):
            self.i += 1
            a = ("bin", "or", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.at("dot", "and"):
            self.i += 1
            a = ("bin", "and", a, self.p_not())
        return a

    def p_not(self):
        if self.at("dot", "not"):
            self.i += 1
            return ("un", "not", self.p_not())
        return self.p_rel()

    _REL = {"eq": "==", "ne": "!=", "lt": "<", "le": "<=", "gt": ">", "ge": ">="}
    _RELOP = {"==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">="}

    def p_rel(self):
        a = self.p_cat()
        t = self.peek()
        if t is not None:
            if t.k == "dot" and t.v in self._REL:
                self.i += 1
                return ("bin", self._REL[t.v], a, self.p_cat())
            if t.k == "op" and t.v in self._RELOP:
                self.i += 1
                return ("bin", self._RELOP[t.v], a, self.p_cat())
        return a

    def p_cat(self):
        a = self.p_add()
        while self.at("op", "//"):
            self.i += 1
            a = ("bin", "//", a, self.p_add())
        return a

    def p_add(self):
        if self.at("op", "-"):
            self.i += 1
            a = ("un", "-", self.p_mul())
        elif self.at("op", "+"):
            self.i += 1
            a = self.p_mul()
        else:
            a = self.p_mul()
        while self.at("op", "+") or self.at("op", "-"):
            op = self.eat("op").v
            a = ("bin", op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.at("op", "*") or self.at("op", "/"):
            op = self.eat("op").v
            a = ("bin", op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.at("op", "**"):
            self.i += 1
            if self.at("op", "-"):
                self.i += 1
                b = ("un", "-", self.p_pow())
            else:
                b = self.p_pow()
            return ("bin", "**", a, b)
        return a

    def p_primary(self):
        t = self.peek()
        if t is None:
            raise SyntaxError(f"unexpected end in {self.t}")
        if t.k == "num":
            self.i += 1
            return ("num", t.v[1], t.v[0])
        if t.k == "str":
            self.i += 1
            return ("str", t.v)
        if t.k == "dot" and t.v in ("true", "false"):
            self.i += 1
            return ("log", t.v == "true")
        if t.k == "op" and t.v == "(":
            self.i += 1
            e = self.expr()
            self.eat("op", ")")
            return ("par", e)
        if t.k == "op" and t.v in ("-", "+"):                 # e.g.  a * -b  (extension) or  .lt. - 1
            self.i += 1
            e = self.p_primary()
            return ("un", "-", e) if t.v == "-" else e
        if t.k == "id":
            return self.p_ref()
        raise SyntaxError(f"unexpected token {t} in {self.t}")

    def p_ref(self):
        chain = []
        while True:
            name = self.eat("id").v
            args = None
            if self.at("op", "("):
                args = self.p_args()
            chain.append((name, args))
            if self.at("op", "%"):
                self.i += 1
                continue
            break
        return ("ref", chain)

    def p_args(self):
        self.eat("op", "(")
        args = []
        if self.at("op", ")"):
            self.i += 1
            return args
        while True:
            args.append(self.p_arg())
            if self.at("op", ","):
                self.i += 1
                continue
            self.eat("op", ")")
            return args

    def p_arg(self):
        # keyword argument
        if self.at("id") and self.peek(1) is not None and self.peek(1).k == "op" and self.peek(1).v == "=":
            name = self.eat("id").v
            self.i += 1
            return ("kw", name, self.expr())
        if self.at("op", "*"):                                 # assumed size / list-directed unit
            self.i += 1
            return ("star",)
        lo = None
        if not self.at("op", ":"):
            lo = self.expr()
            if not self.at("op", ":"):
                return lo
        self.eat("op", ":")
        hi = None
        step = None
        if not (self.done() or self.at("op", ",") or self.at("op", ")") or self.at("op", ":")):
            hi = self.expr()
        if self.at("op", ":"):
            self.i += 1
            step = self.expr()
        return ("range", lo, hi, step)


def parse_expr(text):
    p = Parser(tokenize(text))
    e = p.expr()
    if not p.done():
        raise SyntaxError(f"trailing tokens in expression {text!r}")
    return e


# ================================================================================================== declarations

class Decl:
    __slots__ = ("name", "base", "tname", "rank", "dims", "alloc", "param", "init", "dummy", "modvar")

    def __init__(self, name, base, tname=None, dims=None, alloc=False, param=False, init=None):
        self.name, self.base, self.tname = name, base, tname
        self.dims = dims
        self.rank = len(dims) if dims else 0
        self.alloc, self.param, self.init = alloc, param, init
        self.dummy = False
        self.modvar = False

    def __repr__(self):
        return f"Decl({self.name}:{self.base}{'/' + self.tname if self.tname else ''} rank{self.rank}{' alloc' if self.alloc else ''})"


_TYPEWORDS = ("integer", "real", "double", "doubleprecision", "logical", "character", "type", "complex")


def is_declaration(toks):
    if not toks or toks[0].k != "id" or toks[0].v not in _TYPEWORDS:
        return False
    if toks[0].v == "type":
        return len(toks) > 1 and toks[1].k == "op" and toks[1].v == "("
    if toks[0].v == "double":
        return len(toks) > 1 and toks[1].k == "id" and toks[1].v == "precision"
    # "real = 3" would be an assignment to a variable called real: not used by the reference
    return not (len(toks) > 1 and toks[1].k == "op" and toks[1].v == "=")


def _group(toks, i):
    """toks[i] is '(' -> (tokens inside, index after the matching ')')"""
    depth = 0
    j = i
    while j < len(toks):
        if toks[j].k == "op" and toks[j].v in ("(", "(/"):
            depth += 1
        elif toks[j].k == "op" and toks[j].v in (")", "/)"):
            depth -= 1
            if depth == 0:
                return toks[i + 1:j], j + 1
        j += 1
    raise SyntaxError(f"unbalanced parentheses in {toks}")


def _split_commas(toks):
    parts = []
    cur = []
    depth = 0
    for t in toks:
        if t.k == "op" and t.v in ("(", "(/"):
            depth += 1
        elif t.k == "op" and t.v in (")", "/)"):
            depth -= 1
        if depth == 0 and t.k == "op" and t.v == ",":
            parts.append(cur)
            cur = []
        else:
            cur.append(t)
    parts.append(cur)
    return parts


def _parse_dims(toks):
    dims = []
    for part in _split_commas(toks):
        p = Parser(part)
        dims.append(p.p_arg())
    return dims


def parse_declaration(toks):
    i = 0
    w = toks[0].v
    tname = None
    if w == "double":
        base = "d"
        i = 2
    elif w == "doubleprecision":
        base = "d"
        i = 1
    else:
        base = {"integer": "i", "real": "r", "logical": "l", "character": "c", "type": "t", "complex": "z"}[w]
        i = 1
        if i < len(toks) and toks[i].k == "op" and toks[i].v == "(":
            inner, i = _group(toks, i)
            if base == "t":
                tname = inner[0].v
            elif base == "r" and any((t.k == "num" and t.v[0] == "8") or (t.k == "id" and t.v in ("c_double", "dp", "real64")) for t in inner):
                base = "d"
        elif i < len(toks) and toks[i].k == "op" and toks[i].v == "*":        # real*8, character*20
            if base == "r" and toks[i + 1].k == "num" and toks[i + 1].v[0] == "8":
                base = "d"
            i += 2
    dims_attr = None
    alloc = False
    param = False
    while i < len(toks) and toks[i].k == "op" and toks[i].v == ",":
        i += 1
        attr = toks[i].v
        i += 1
        grp = None
        if i < len(toks) and toks[i].k == "op" and toks[i].v == "(":
            grp, i = _group(toks, i)
        if attr == "dimension":
            dims_attr = _parse_dims(grp)
        elif attr == "allocatable":
            alloc = True
        elif attr == "parameter":
            param = True
    if i < len(toks) and toks[i].k == "op" and toks[i].v == "::":
        i += 1
    out = []
    for ent in _split_commas(toks[i:]):
        if not ent:
            continue
        name = ent[0].v
        j = 1
        dims = dims_attr
        init = None
        if j < len(ent) and ent[j].k == "op" and ent[j].v == "(":
            grp, j = _group(ent, j)
            dims = _parse_dims(grp)
        if j < len(ent) and ent[j].k == "op" and ent[j].v == "*":              # character name*20
            j += 2
        if j < len(ent) and ent[j].k == "op" and ent[j].v == "=":
            p = Parser(ent, j + 1)
            init = p.expr()
        out.append(Decl(name, base, tname, dims, alloc, param, init))
    return out


# ================================================================================================== program structure

class TypeDef:
    def __init__(self, name):
        self.name = name
        self.comps = {}          # name -> Decl (ordered)


class Proc:
    def __init__(self, kind, name, args, result, module, src):
        self.kind, self.name, self.args, self.result, self.module, self.src = kind, name, args, result, module, src
        self.decls = {}
        self.body = []           # [(line, label, text)]
        self.copy_out = []       # positions of dummies returned to the caller
        self.uses = []
        self.broken = None
        self.host = None         # the program / procedure an internal procedure is contained in
        self.formats = {}        # statement label -> format text


class Program:
    """everything parsed from a set of source files, in one flat name space (module procedure names are unique in
    the reference)"""

    def __init__(self):
        self.types = {}
        self.procs = {}
        self.generics = {}       # generic name -> [specific names]
        self.modvars = {}        # name -> Decl (parameters and module variables)
        self.modvar_order = []
        self.warnings = []
        self.externals = set()   # procedures declared in interface bodies (defined elsewhere, e.g. bind(C))
        for name, val in (("c_int", 4), ("c_int32_t", 4), ("c_int64_t", 8), ("c_double", 8), ("c_size_t", 8), ("c_null_ptr", 0)):
            d = Decl(name, "i", param=True, init=("num", "i", str(val)))      # the named constants of iso_c_binding
            d.modvar = True
            self.modvars[name] = d
            self.modvar_order.append(name)
        self.types["c_ptr"] = TypeDef("c_ptr")
        self.linemap = {}        # line of the generated text -> (procedure, "file.f90:line") of the reference statement

    # ---------------------------------------------------------------------------------------------- parsing
    def add_source(self, text, src="<text>"):
        lines = logical_lines(text)
        module = None
        cur_type = None
        cur_proc = None
        cur_iface = None
        in_exec = False
        host = None
        for ln, s in lines:
            label = None
            m = re.match(r"^(\d+)\s+(.*)$", s)
            if m:
                label, s = m.group(1), m.group(2)
            if cur_type is not None:
                if re.match(r"^end\s*type\b", s):
                    cur_type = None
                    continue
                try:
                    toks = tokenize(s)
                    if is_declaration(toks):
                        for d in parse_declaration(toks):
                            cur_type.comps[d.name] = d
                except SyntaxError as e:
                    self.warnings.append(f"{src}:{ln}: component outside the subset: {e}")
                continue
            if cur_iface is not None:
                if re.match(r"^end\s*interface\b", s):
                    cur_iface = None
                    continue
                m = re.match(r"^module\s+procedure\s+(.*)$", s)
                if m and cur_iface:
                    self.generics.setdefault(cur_iface, []).extend(x.strip() for x in m.group(1).split(","))
                m = re.match(r"^(?:[\w()]+\s+)?(function|subroutine)\s+(\w+)", s)
                if m and not s.startswith("end"):
                    self.externals.add(m.group(2))
                continue
            if cur_proc is not None:
                if re.match(r"^end\s*(subroutine|function|program)\b", s) or s == "end":
                    self.procs[cur_proc.name] = cur_proc
                    cur_proc = None
                    continue
                if s == "contains":                                # internal procedures follow: they see the host's entities
                    self.procs[cur_proc.name] = cur_proc
                    host = cur_proc
                    cur_proc = None
                    continue
                if label is not None and re.match(r"^format\s*\(", s):
                    cur_proc.formats[label] = s[s.index("("):]
                    continue
                if not in_exec:
                    if re.match(r"^(implicit|private|public|intrinsic|external|save)\b", s):
                        continue
                    m = re.match(r"^parameter\s*\((.*)\)$", s)
                    if m:
                        try:
                            for part in _split_commas(tokenize(m.group(1))):
                                d = cur_proc.decls[part[0].v]
                                d.param = True
                                d.init = Parser(part, 2).expr()
                        except (SyntaxError, KeyError):
                            cur_proc.broken = f"parameter statement outside the subset: {s[:60]}"
                        continue
                    m = re.match(r"^use\b[\s,]*(?:(?:non_)?intrinsic\s*::\s*)?(\w+)", s)
                    if m:
                        cur_proc.uses.append(m.group(1))
                        continue
                    try:
                        toks = tokenize(s)
                        if is_declaration(toks):
                            for d in parse_declaration(toks):
                                cur_proc.decls[d.name] = d
                            continue
                    except SyntaxError as e:
                        cur_proc.broken = f"declaration outside the subset: {s[:60]}"
                        continue
                    in_exec = True
                cur_proc.body.append((ln, label, s))
                continue
            # module level
            m = re.match(r"^module\s+(\w+)$", s)
            if m and not s.startswith("module procedure"):
                module = m.group(1)
                continue
            if re.match(r"^end\s*(module|program)?\b", s) and not re.match(r"^end\s*(if|do)", s):
                host = None
                continue
            m = re.match(r"^program\s+(\w+)$", s)
            if m:                                                  # a main program: a procedure without arguments
                cur_proc = Proc("sub", m.group(1), [], None, module, src)
                in_exec = False
                continue
            if s == "contains" or re.match(r"^(implicit|private|public|use)\b", s):
                continue
            m = re.match(r"^interface\s*(\w+)?$", s)
            if m:
                cur_iface = m.group(1) or ""
                continue
            m = re.match(r"^type\s*(?:,\s*[\w()]+\s*)*(?:::)?\s*(\w+)$", s)
            if m and not s.startswith("type("):
                cur_type = TypeDef(m.group(1))
                self.types[cur_type.name] = cur_type
                continue
            m = re.match(r"^(?:recursive\s+)?subroutine\s+(\w+)\s*(?:\((.*)\))?$", s)
            if m:
                args = [a.strip() for a in (m.group(2) or "").split(",") if a.strip()]
                cur_proc = Proc("sub", m.group(1), args, None, module, src)
                cur_proc.host = host
                in_exec = False
                continue
            m = re.match(r"^(?:recursive\s+)?(?:[\w\s()*=]+?\s+)?function\s+(\w+)\s*\((.*?)\)\s*(?:result\s*\(\s*(\w+)\s*\))?$", s)
            if m:
                args = [a.strip() for a in m.group(2).split(",") if a.strip()]
                cur_proc = Proc("fun", m.group(1), args, m.group(3) or m.group(1), module, src)
                cur_proc.host = host
                in_exec = False
                continue
            try:
                toks = tokenize(s)
                if is_declaration(toks):
                    for d in parse_declaration(toks):
                        d.modvar = True
                        self.modvars[d.name] = d
                        self.modvar_order.append(d.name)
                    continue
            except SyntaxError as e:
                self.warnings.append(f"{src}:{ln}: module entity outside the subset: {s[:60]}")
                continue
            self.warnings.append(f"{src}:{ln}: ignored at module level: {s}")

    def add_file(self, path):
        with open(path, "r", errors="replace") as f:
            self.add_source(f.read(), path)

    # ---------------------------------------------------------------------------------------------- code generation
    def finish(self):
        for p in self.procs.values():
            for k, a in enumerate(p.args):
                d = p.decls.get(a)
                if d is None:
                    d = Decl(a, "i" if a[0] in "ijklmn" else "r")
                    p.decls[a] = d
                    self.warnings.append(f"{p.name}: dummy {a} typed implicitly")
                d.dummy = True
            if p.kind == "sub":
                p.copy_out = [k for k, a in enumerate(p.args)
                              if (p.decls[a].rank == 0 and p.decls[a].base != "t") or p.decls[a].alloc]

    def generate(self, only=None):
        """Python source of every type, module entity and procedure (procedures outside the subset become stubs)"""
        self.finish()
        out = ["_M = Namespace()", "_NOT_TRANSLATED = {}", ""]
        for t in self.types.values():
            out += self._gen_type(t)
        g = Gen(self, None)
        for name in self.modvar_order:
            d = self.modvars[name]
            try:
                out.append(f"_M.{pyname(name)} = {g.initial_value(d)}")
            except (SyntaxError, NotImplementedError, KeyError) as e:
                self.warnings.append(f"module entity {name} outside the subset: {e}")
        out.append("")
        for p in self.procs.values():
            if only is not None and p.name not in only:
                continue
            try:
                if p.broken:
                    raise NotImplementedError(p.broken)
                g = Gen(self, p)
                lines = g.gen_proc()
                for k, w in enumerate(g.origin):
                    if w:
                        self.linemap[len(out) + k + 1] = (p.name, w)
                out += lines
            except (SyntaxError, NotImplementedError, KeyError, AssertionError) as e:
                msg = f"{type(e).__name__}: {e}"
                out.append(f"def f_{p.name}(*a, **k):")
                out.append(f"    raise FUnsupported({('procedure ' + p.name + ' is outside the translated subset: ' + msg)!r})")
                out.append(f"_NOT_TRANSLATED[{p.name!r}] = {msg!r}")
            out.append("")
        return "\n".join(out)

    def _gen_type(self, t):
        g = Gen(self, None)
        names = [pyname(c) for c in t.comps]
        out = [f"class T_{t.name}:", f"    __slots__ = {tuple(names)!r}", "    def __init__(self):"]
        if not names:
            out.append("        pass")
        for c, d in t.comps.items():
            out.append(f"        self.{pyname(c)} = {g.initial_value(d)}")
        out.append("")
        return out

    def load(self, only=None, extern=None):
        """exec the generated text -> name space dict (procedures are f_<name>, types T_<name>, module entities _M.<name>)"""
        import os
        import sys
        here = os.path.dirname(os.path.abspath(__file__))
        if here not in sys.path:
            sys.path.insert(0, here)
        import f90_runtime
        ns = {k: v for k, v in vars(f90_runtime).items() if not k.startswith("__")}
        ns["__runtime__"] = f90_runtime
        code = self.generate(only)
        exec(compile(code, "<f90_translate>", "exec"), ns)
        ns["__source__"] = code
        ns["__linemap__"] = dict(self.linemap)
        if extern:
            ns.update(extern)
        return ns


def pyname(n):
    return n + "_" if keyword.iskeyword(n) or n in ("None", "True", "False") else n


_CV = {"i": "cv_i", "r": "cv_r", "d": "cv_d", "l": "cv_l", "c": "cv_c", "t": "cv_t", "z": "cv_c"}

_INTRINSICS = {"exp", "log", "cos", "sin", "tan", "atan", "sqrt", "abs", "min", "max", "dble", "real", "int", "nint", "floor",
               "ceiling", "mod", "sign", "size", "allocated", "sum", "minval", "maxval", "cshift", "trim", "len_trim",
               "adjustl", "huge", "isnan", "count", "len", "index", "iachar", "ichar", "achar", "char",
               "c_loc", "c_associated", "c_sizeof", "command_argument_count", "iand", "ior", "ieor", "ishft", "btest"}


class Gen:
    """code generator for one procedure"""

    def __init__(self, prog, proc):
        self.prog = prog
        self.p = proc
        self.lines = []
        self.origin = []         # per emitted line: "file.f90:line" of the statement it came from ("" for prologue lines)
        self.ind = 1
        self.where = ""

    # ---------------------------------------------------------------------------------------------- symbols
    def lookup(self, name):
        if self.p is not None and name in self.p.decls:
            return self.p.decls[name]
        if self.p is not None and self.p.host is not None and name in self.p.host.decls:
            raise NotImplementedError(f"host association of {name}")      # parsed, not executed

        if name in self.prog.modvars:
            return self.prog.modvars[name]
        return None

    def var_code(self, d):
        return f"_M.{pyname(d.name)}" if d.modvar else f"v_{d.name}"

    def initial_value(self, d):
        if d.init is not None and (d.param or d.rank == 0):
            return f"{_CV[d.base]}({self.expr(d.init)})"
        if d.alloc:
            return "None"
        if d.rank:
            if any(x == ("star",) or (isinstance(x, tuple) and x[0] == "range" and x[2] is None) for x in d.dims):
                return "None"
            return self.alloc_code(d, d.dims)
        if d.base == "t":
            return f"T_{d.tname}()"
        return {"i": "0", "r": "F32(0.0)", "d": "D0", "l": "False", "c": "''", "z": "0j"}[d.base]

    def alloc_code(self, d, dims):
        b = []
        for x in dims:
            if isinstance(x, tuple) and x[0] == "range":
                b.append(f"({self.expr(x[1]) if x[1] is not None else 1}, {self.expr(x[2])})")
            else:
                b.append(f"(1, {self.expr(x)})")
        fac = f", T_{d.tname}" if d.base == "t" else ""
        return f"f_alloc({d.base!r}, [{', '.join(b)}]{fac})"

    # ---------------------------------------------------------------------------------------------- expressions
    def expr(self, e):
        k = e[0]
        if k == "num":
            txt = e[2]
            if e[1] == "i":
                return str(int(txt))
            if e[1] == "d":
                return f"F64({float(txt.replace('d', 'e'))!r})"
            return f"F32({txt.replace('d', 'e')!r})"
        if k == "str":
            return repr(e[1])
        if k == "log":
            return "True" if e[1] else "False"
        if k == "par":
            return f"({self.expr(e[1])})"
        if k == "un":
            if e[1] == "not":
                return f"(not {self.expr(e[2])})"
            return f"(-{self.expr(e[2])})"
        if k == "bin":
            op, a, b = e[1], self.expr(e[2]), self.expr(e[3])
            if op == "/":
                return f"f_div({a}, {b})"
            if op == "**":
                return f"f_pow({a}, {b})"
            if op == "//":
                return f"(str({a}) + str({b}))"
            if op == "eqv":
                return f"(bool({a}) == bool({b}))"
            if op == "neqv":
                return f"(bool({a}) != bool({b}))"
            return f"({a} {op} {b})"
        if k == "ref":
            return self.ref(e[1])[0]
        if k == "range":
            return self.sub(e)
        raise NotImplementedError(f"expression node {k}")

    def sub(self, a):
        if isinstance(a, tuple) and a[0] == "range":
            lo = self.expr(a[1]) if a[1] is not None else "None"
            hi = self.expr(a[2]) if a[2] is not None else "None"
            st = self.expr(a[3]) if a[3] is not None else "None"
            return f"f_slice({lo}, {hi}, {st})"
        if isinstance(a, tuple) and a[0] in ("kw", "star"):
            raise NotImplementedError("keyword / * subscript")
        return self.expr(a)

    def ref(self, chain, lvalue=False):
        """-> (code, Decl-like description of the result: base, tname, rank, kind)
        kind: 'scalar' | 'array' (whole) | 'section' | 'gather' (component through an array of derived type) | 'call'"""
        name, args = chain[0]
        d = self.lookup(name)
        if d is None:
            if lvalue:
                raise KeyError(f"assignment to undeclared name {name}")
            # function reference
            if args is None:
                raise KeyError(f"undeclared name {name}")
            if len(chain) != 1:
                raise NotImplementedError("component of a function result")
            if name in self.prog.procs and self.prog.procs[name].kind == "fun":
                f = self.prog.procs[name]
                rd = f.decls.get(f.result)
                return (f"f_{name}({', '.join(self.expr(a) for a in args)})",
                        (rd.base if rd else "d", rd.tname if rd else None, rd.rank if rd else 0, "call"))
            if name in self.prog.externals:
                return (f"f_{name}({', '.join(self.expr(a) for a in args)})", (None, None, 0, "call"))
            if name in _INTRINSICS:
                pa = []
                for a in args:
                    if isinstance(a, tuple) and a[0] == "kw":
                        pa.append(f"{a[1]}={self.expr(a[2])}")
                    else:
                        pa.append(self.expr(a))
                return (f"i_{name}({', '.join(pa)})", (None, None, 0, "call"))
            raise KeyError(f"unknown function {name}")
        code = self.var_code(d)
        base, tname, rank = d.base, d.tname, d.rank
        kind = "array" if rank else "scalar"
        gather_base = None
        for idx, (nm, ar) in enumerate(chain):
            if idx > 0:
                comp = self.prog.types[tname].comps[nm]
                if rank:                                       # x%c with x an array: gather / scatter
                    if gather_base is not None or comp.rank:
                        raise NotImplementedError("nested array component reference")
                    gather_base = (code, nm)
                    kind = "gather"
                    base, tname = comp.base, comp.tname
                else:
                    code = f"{code}.{pyname(nm)}"
                    base, tname, rank = comp.base, comp.tname, comp.rank
                    kind = "array" if rank else "scalar"
            if ar is not None:
                if kind == "gather":
                    raise NotImplementedError("subscript after array component reference")
                if not rank:
                    if base == "c" and len(ar) == 1 and isinstance(ar[0], tuple) and ar[0][0] == "range" and idx == len(chain) - 1:
                        lo = self.expr(ar[0][1]) if ar[0][1] is not None else "None"
                        hi = self.expr(ar[0][2]) if ar[0][2] is not None else "None"
                        if lvalue:
                            return (code, lo, hi), (base, tname, 0, "substr")
                        return f"f_substr({code}, {lo}, {hi})", (base, tname, 0, "scalar")
                    if base == "c":
                        raise NotImplementedError("substring of an array element")
                    raise SyntaxError(f"subscript on scalar {nm}")
                subs = [self.sub(a) for a in ar]
                nsec = sum(1 for a in ar if isinstance(a, tuple) and a[0] == "range")
                code = f"{code}[{', '.join(subs)}]"
                rank = nsec
                kind = "section" if nsec else "scalar"
        if kind == "gather":
            if not lvalue:
                code = f"f_gather({gather_base[0]}, {pyname(gather_base[1])!r})"
            else:
                code = gather_base
        return code, (base, tname, rank, kind)

    # ---------------------------------------------------------------------------------------------- statements
    def emit(self, s):
        self.lines.append("    " * self.ind + s)
        self.origin.append(self.where)

    def gen_proc(self):
        p = self.p
        args = ", ".join(f"v_{a}" for a in p.args)
        self.lines = [f"def f_{p.name}({args}):"]
        self.origin = [""]
        self.ind = 1
        # locals
        for d in p.decls.values():
            if d.dummy:
                continue
            if d.param:
                self.emit(f"v_{d.name} = {self.initial_value(d)}")
        for d in p.decls.values():
            if d.dummy or d.param:
                continue
            self.emit(f"v_{d.name} = {self.initial_value(d)}")
        if p.kind == "fun" and p.result not in p.decls:
            raise NotImplementedError(f"function {p.name}: result typed in the function statement")
        self.ret = (f"return ({', '.join('v_' + p.args[k] for k in p.copy_out)}{',' if len(p.copy_out) == 1 else ''})"
                    if p.kind == "sub" else f"return v_{p.result}")
        stack = []                                             # ('if',) | ('do', label)
        for ln, label, s in p.body:
            self.where = f"{p.src.split('/')[-1]}:{ln}"
            self.stmt(s, stack, label)
            while label is not None and stack and stack[-1][0] == "do" and stack[-1][1] == label:
                stack.pop()
                self.ind -= 1
        if stack:
            raise SyntaxError(f"{p.name}: unclosed block {stack}")
        self.where = ""
        self.emit(self.ret)
        return self.lines

    def io(self, toks):
        """print / write: the first string literal of the list names the event, the values of the other list items are
        recorded (arrays as copies); a list outside the subset (implied DO, array constructor) is recorded without values"""
        internal = None
        if toks[0].v == "write":
            grp, j = _group(toks, 1)
            items = toks[j:]
            ctl = _split_commas(grp)
            first = ctl[0]
            if first and first[0].k == "id" and len(ctl) > 1:
                d = self.lookup(first[0].v)
                if d is not None and d.base in ("c", "t"):
                    try:
                        code, (base, _, rank, kind) = self.ref(Parser(first).p_ref()[1], lvalue=True)
                        if base == "c" and kind == "scalar":
                            a = Parser(ctl[1]).p_arg()
                            if not (isinstance(a, tuple) and a[0] in ("star", "kw", "num")):
                                internal = (code, self.expr(a))
                    except (SyntaxError, NotImplementedError, KeyError):
                        internal = None
        else:
            parts = _split_commas(toks[1:])
            items = [t for part in parts[1:] for t in part + [Tok("op", ",", 0, 0)]][:-1] if len(parts) > 1 else []
        txt = next((t.v for t in items if t.k == "str" and t.v.strip()), "")
        vals = []
        try:
            for part in _split_commas(items) if items else []:
                if not part or (len(part) == 1 and part[0].k == "str"):
                    continue
                pr = Parser(part)
                e = pr.expr()
                if not pr.done():
                    raise SyntaxError("io item")
                vals.append(self.expr(e))
        except (SyntaxError, NotImplementedError, KeyError):
            vals = []
        if internal is not None and (vals or not items):
            # WRITE to a character variable: formatted for the integer / character descriptors (names of output files, headers)
            allv = []
            for part in _split_commas(items) if items else []:
                if part:
                    allv.append(self.expr(Parser(part).expr()))
            self.emit(f"{internal[0]} = f_fmt({self.where!r}, {internal[1]}, [{', '.join(allv)}])")
            return
        self.emit(f"f_io({self.where!r}, {txt!r}, [{', '.join(vals)}])")

    def read(self, toks):
        """READ (unit, fmt [, iostat = v]) items -- list-directed or '(A)'; items: variables, array elements, whole arrays,
        sections, one-level implied DO"""
        grp, j = _group(toks, 1)
        ctl = _split_commas(grp)
        u0 = Parser(ctl[0]).p_arg()
        if u0 == ("star",):
            self.emit(f"f_unsupported({self.where!r}, 'read from standard input')")
            return
        unit = self.expr(u0)
        fmt = "*"
        iostat = None
        for k, part in enumerate(ctl[1:]):
            a = Parser(part).p_arg()
            if isinstance(a, tuple) and a[0] == "kw":
                if a[1] == "iostat":
                    iostat = self.ref(a[2][1], lvalue=True)[0]
                elif a[1] == "fmt":
                    a = a[2]
                else:
                    raise NotImplementedError(f"read control {a[1]}")
            if k == 0 and not (isinstance(a, tuple) and a[0] == "kw"):
                if a == ("star",):
                    fmt = "*"
                elif a[0] == "str":
                    fmt = a[1]
                elif a[0] == "num" and a[2] in self.p.formats:
                    fmt = self.p.formats[a[2]]
                else:
                    raise NotImplementedError("read with a format held in a variable")
        spec, stores = [], []
        items = _split_commas(toks[j:]) if j < len(toks) else []
        for part in items:
            if not part:
                continue
            n = len(spec)
            if part[0].k == "op" and part[0].v == "(" and any(t.k == "op" and t.v == "=" for t in part):
                inner, _ = _group(part, 0)                          # implied DO: (a(i), b(i), i = lo, hi)
                pieces = _split_commas(inner)
                eq = next(i for i, pc in enumerate(pieces) if len(pc) > 1 and pc[1].k == "op" and pc[1].v == "=")
                var = self.lookup(pieces[eq][0].v)
                lo = self.expr(Parser(pieces[eq][2:]).expr())
                hi = self.expr(Parser(pieces[eq + 1]).expr())
                targets = [self.ref(Parser(pc).p_ref()[1], lvalue=True) for pc in pieces[:eq]]
                bases = {t[1][0] for t in targets}
                if len(bases) != 1 or any(t[1][3] != "scalar" for t in targets):
                    raise NotImplementedError("implied DO over mixed or array items")
                base = bases.pop()
                spec.append(f"({base!r}, len(f_range({lo}, {hi})) * {len(targets)})")
                body = [f"_it = iter(_v[{n}])", f"for {self.var_code(var)} in f_range({lo}, {hi}):"]
                for code, (b, tn, rk, kd) in targets:
                    body.append(f"    {code} = " + (f"next(_it)" if code.endswith("]") else f"{_CV[b]}(next(_it))"))
                stores.append(body)
                continue
            code, (base, tname, rank, kind) = self.ref(Parser(part).p_ref()[1], lvalue=True)
            if kind == "scalar":
                spec.append(f"({base!r}, None)")
                stores.append([f"{code} = _v[{n}]" if code.endswith("]") else f"{code} = {_CV[base]}(_v[{n}])"])
            elif kind in ("array", "section"):
                spec.append(f"({base!r}, i_size({code}))")
                stores.append([f"f_store({code}, _v[{n}], {base!r})"])
            else:
                raise NotImplementedError("read into a component of an array of derived type")
        self.emit(f"_io, _v = f_read({self.where!r}, {unit}, {fmt!r}, [{', '.join(spec)}], {iostat is not None})")
        if iostat is not None:
            self.emit(f"{iostat} = _io")
            self.emit("if _io == 0:")
            self.ind += 1
        for body in stores:
            for line in body:
                self.emit(line)
        if iostat is not None:
            if not stores:
                self.emit("pass")
            self.ind -= 1

    def stmt(self, s, stack, label=None):
        try:
            toks = tokenize(s)
        except SyntaxError:
            if re.match(r"^(print|write)\b", s):
                self.emit(f"f_io({self.where!r}, '', [])")
                return
            raise
        t0 = toks[0]
        w = t0.v if t0.k == "id" else None
        nxt = toks[1] if len(toks) > 1 else None
        is_assign = self._top_level_assign(toks)
        if w == "end" or (w in ("endif", "enddo") and len(toks) == 1):
            what = w[3:] if w != "end" else (toks[1].v if len(toks) > 1 else "")
            if what == "if":
                assert stack and stack[-1][0] == "if", f"{self.where}: end if without if"
                stack.pop()
                self.ind -= 1
                return
            if what == "do":
                assert stack and stack[-1][0] == "do", f"{self.where}: end do without do"
                if stack[-1][1] is not None and stack[-1][1] == label:
                    return                                     # closed by the label logic in gen_proc
                stack.pop()
                self.ind -= 1
                return
            if what == "select":
                assert stack and stack[-1][0] == "select", f"{self.where}: end select without select"
                if stack[-1][2]:
                    self.ind -= 1
                stack.pop()
                return
            raise NotImplementedError(f"end {what}")
        if is_assign is None:
            if w == "if" and nxt is not None and nxt.v == "(":
                grp, j = _group(toks, 1)
                cond = self.expr(Parser(grp).expr())
                if j < len(toks) and toks[j].k == "id" and toks[j].v == "then" and j == len(toks) - 1:
                    self.emit(f"if {cond}:")
                    stack.append(("if",))
                    self.ind += 1
                    self.emit("pass")
                    return
                self.emit(f"if {cond}:")
                self.ind += 1
                self.stmt(s[toks[j].pos:], stack)
                self.ind -= 1
                return
            if w in ("else", "elseif"):
                assert stack and stack[-1][0] == "if", f"{self.where}: else without if"
                j = 1
                if w == "else" and nxt is not None and nxt.k == "id" and nxt.v == "if":
                    j = 2
                if w == "elseif" or j == 2:
                    grp, _ = _group(toks, j)
                    cond = self.expr(Parser(grp).expr())
                    self.ind -= 1
                    self.emit(f"elif {cond}:")
                else:
                    self.ind -= 1
                    self.emit("else:")
                self.ind += 1
                self.emit("pass")
                return
            if w == "do":
                j = 1
                lab = None
                if nxt is not None and nxt.k == "num":
                    lab = nxt.v[0]
                    j = 2
                if j >= len(toks):
                    self.emit("while True:")
                elif toks[j].k == "id" and toks[j].v == "while":
                    grp, _ = _group(toks, j + 1)
                    self.emit(f"while {self.expr(Parser(grp).expr())}:")
                else:
                    var = toks[j].v
                    d = self.lookup(var)
                    if d is None:
                        raise KeyError(f"undeclared loop variable {var}")
                    parts = _split_commas(toks[j + 2:])
                    ex = [self.expr(Parser(x).expr()) for x in parts]
                    self.emit(f"for {self.var_code(d)} in f_range({', '.join(ex)}):")
                stack.append(("do", lab))
                self.ind += 1
                self.emit("pass")
                return
            if w == "continue":
                self.emit("pass")
                return
            if w == "exit":
                self.emit("break")
                return
            if w == "cycle":
                self.emit("continue")
                return
            if w == "return":
                self.emit(self.ret)
                return
            if w == "stop":
                txt = toks[1].v if len(toks) > 1 and toks[1].k == "str" else ""
                self.emit(f"f_stop({self.where!r}, {txt!r})")
                return
            if w == "call":
                self.call(toks)
                return
            if w == "allocate":
                grp, _ = _group(toks, 1)
                for part in _split_commas(grp):
                    a = Parser(part).p_arg()
                    if a[0] == "kw":
                        if a[1] == "stat":
                            code, _ = self.ref(a[2][1], lvalue=True)
                            self.emit(f"{code} = 0")
                        continue
                    chain = a[1]
                    dims = chain[-1][1]
                    bare = chain[:-1] + [(chain[-1][0], None)]
                    code, (base, tname, rank, kind) = self.ref(bare, lvalue=True)
                    fake = Decl("", base, tname)
                    self.emit(f"{code} = {self.alloc_code(fake, dims)}")
                return
            if w == "deallocate":
                grp, _ = _group(toks, 1)
                for part in _split_commas(grp):
                    a = Parser(part).p_arg()
                    if a[0] == "kw":
                        continue
                    code, _ = self.ref(a[1], lvalue=True)
                    self.emit(f"{code} = None")
                return
            if w in ("print", "write"):
                self.io(toks)
                return
            if w == "select" and nxt is not None and nxt.v == "case":
                grp, _ = _group(toks, 2)
                self.sel_n = getattr(self, "sel_n", 0) + 1
                var = f"_sel{self.sel_n}"
                self.emit(f"{var} = {self.expr(Parser(grp).expr())}")
                stack.append(["select", var, False])
                return
            if w == "case":
                assert stack and stack[-1][0] == "select", f"{self.where}: case outside select"
                top = stack[-1]
                if top[2]:
                    self.ind -= 1
                if nxt is not None and nxt.k == "id" and nxt.v == "default":
                    self.emit("else:" if top[2] else "if True:")
                else:
                    grp, _ = _group(toks, 1)
                    vals = []
                    for part in _split_commas(grp):
                        a = Parser(part).p_arg()
                        if isinstance(a, tuple) and a[0] == "range":
                            raise NotImplementedError("case range")
                        vals.append(self.expr(a))
                    self.emit(f"{'elif' if top[2] else 'if'} f_case({top[1]}, {', '.join(vals)}):")
                top[2] = True
                self.ind += 1
                self.emit("pass")
                return
            if w == "read":
                self.read(toks)
                return
            if w == "open":
                grp, _ = _group(toks, 1)
                pos, kws = [], []
                ios = None
                for part in _split_commas(grp):
                    a = Parser(part).p_arg()
                    if isinstance(a, tuple) and a[0] == "kw":
                        if a[1] in ("iostat",):
                            ios, _ = self.ref(a[2][1], lvalue=True)
                            kws.append("iostat=True")
                        elif a[1] == "unit":
                            pos.append(self.expr(a[2]))
                        else:
                            kws.append(f"{a[1]}={self.expr(a[2])}")
                    else:
                        pos.append(self.expr(a))
                self.emit(f"{ios + ' = ' if ios else ''}f_open({self.where!r}, {pos[0]}{''.join(', ' + k for k in kws)})")
                return
            if w in ("close", "rewind", "backspace"):
                if nxt is not None and nxt.k == "op" and nxt.v == "(":
                    grp, _ = _group(toks, 1)
                    a = Parser(_split_commas(grp)[0]).p_arg()
                    u = self.expr(a[2] if isinstance(a, tuple) and a[0] == "kw" else a)
                else:
                    u = self.expr(Parser(toks, 1).expr())
                self.emit(f"f_{w}({u})")
                return
            if w in ("inquire", "format", "flush"):
                self.emit(f"f_unsupported({self.where!r}, {w!r})")
                return
            raise NotImplementedError(f"statement {s!r}")
        # assignment
        lhs = Parser(toks[:is_assign]).p_ref()
        pr = Parser(toks, is_assign + 1)
        rhs = pr.expr()
        if not pr.done():
            raise SyntaxError(f"trailing tokens in {s!r}")
        code, (base, tname, rank, kind) = self.ref(lhs[1], lvalue=True)
        r = self.expr(rhs)
        if kind == "substr":
            self.emit(f"{code[0]} = f_setsub({code[0]}, {code[1]}, {code[2]}, {r})")
        elif kind == "gather":
            self.emit(f"f_scatter({code[0]}, {pyname(code[1])!r}, {r}, {_CV[base]})")
        elif kind == "array":
            self.emit(f"{code} = f_aassign({code}, {r})")
        elif kind == "section":
            self.emit(f"{code} = {r}")
        elif code.endswith("]"):
            if base == "t":
                self.emit(f"{code} = cv_t({r})")
            else:
                self.emit(f"{code} = {r}")
        elif base == "t" and tname == "c_ptr":
            self.emit(f"{code} = {r}")                               # a C address: the reference to the object, not a copy
        else:
            self.emit(f"{code} = {_CV[base]}({r})")

    @staticmethod
    def _top_level_assign(toks):
        depth = 0
        for i, t in enumerate(toks):
            if t.k == "op":
                if t.v in ("(", "(/"):
                    depth += 1
                elif t.v in (")", "/)"):
                    depth -= 1
                elif t.v == "=" and depth == 0:
                    # "if (...) x = 1" is an IF statement, not an assignment to if(...)
                    if toks[0].k == "id" and toks[0].v in ("if", "do", "else", "elseif") and i > 0:
                        if toks[0].v == "if" and toks[1].k == "op" and toks[1].v == "(":
                            _, j = _group(toks, 1)
                            if j < i:
                                return None
                        elif toks[0].v == "do":
                            return None
                    return i
        return None

    def call(self, toks):
        name = toks[1].v
        args = []
        if len(toks) > 2:
            args = Parser(toks, 2).p_args()
        if name == "get_command_argument":
            dst, _ = self.ref(args[1][1], lvalue=True)
            self.emit(f"{dst} = f_get_command_argument({self.expr(args[0])})")
            return
        if name == "move_alloc":
            src, _ = self.ref(args[0][1], lvalue=True)
            dst, _ = self.ref(args[1][1], lvalue=True)
            self.emit(f"{dst} = {src}")
            self.emit(f"{src} = None")
            return
        target = name
        if name in self.prog.generics:
            target = self.resolve_generic(name, args)
        callee = self.prog.procs.get(target)
        if any(isinstance(a, tuple) and a[0] == "kw" for a in args):
            raise NotImplementedError("keyword arguments in a call")
        pa = [self.expr(a) for a in args]
        if callee is None:
            self.emit(f"f_{target}({', '.join(pa)})")          # external hook supplied by the driver (or NameError)
            return
        if len(pa) != len(callee.args):
            raise NotImplementedError(f"call {name}: {len(pa)} actual vs {len(callee.args)} dummy arguments")
        backs = []
        for pos, k in enumerate(callee.copy_out):
            a = args[k]
            if a[0] != "ref":
                continue
            d = self.lookup(a[1][0][0])
            if d is None or d.param:
                continue
            if any(ar is not None and any(isinstance(x, tuple) and x[0] == "range" for x in ar) for _, ar in a[1]):
                continue
            code, (base, tname, rank, kind) = self.ref(a[1], lvalue=True)
            if kind in ("gather", "call"):
                continue
            backs.append((code, pos))
        if not callee.copy_out:
            self.emit(f"f_{target}({', '.join(pa)})")
            return
        self.emit(f"_r = f_{target}({', '.join(pa)})")
        for code, pos in backs:
            self.emit(f"{code} = _r[{pos}]")

    def resolve_generic(self, name, args):
        a0 = args[0]
        base = rank = None
        if a0[0] == "ref" and self.lookup(a0[1][0][0]) is not None:
            _, (base, _, rank, _) = self.ref(a0[1], lvalue=True)
        for cand in self.prog.generics[name]:
            c = self.prog.procs.get(cand)
            if c is None or len(c.args) != len(args):
                continue
            cd = c.decls.get(c.args[0])
            if base is None or (cd.base == base and cd.rank == rank):
                return cand
        raise KeyError(f"no specific procedure of {name} matches")
