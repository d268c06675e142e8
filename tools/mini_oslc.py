#!/usr/bin/env python3
"""mini_oslc — a small OSL -> .oso (OSO 1.00) compiler used ONLY to produce
test fixtures.

The north star keeps the real front end (oslc / liboslcomp) as it is; oslc
cannot be built in this environment (no flex/bison/LLVM/OIIO), and the
reference ships no precompiled .oso.  This tool compiles the subset of OSL the
testsuite shaders on the hot path use, so the oracle and the CUDA back end can
be driven by the reference's own shaders and checked against its goldens.

Behaviour follows the reference compiler where it affects the emitted ops:
  * .oso layout                       src/liboslcomp/oslcomp.cpp:743-999
  * overload scoring                  src/liboslcomp/typecheck.cpp:1438-1760
  * expected-type propagation         src/liboslcomp/typecheck.cpp:20-50,466
  * op emission (if/loops/logic/?:)   src/liboslcomp/codegen.cpp:1378-1655
  * user functions inlined under a `functioncall` op with formals aliased to
    actuals                           src/liboslcomp/codegen.cpp:1796-1920
  * builtin declarations come from the reference's own stdosl.h, parsed at
    fixture-generation time (it is not copied into this repo).
Not supported: structs, closure keyword args, operator overloading.
"""
import os
import re
import struct
import sys

# ----------------------------------------------------------------------------
# types
# ----------------------------------------------------------------------------
TRIPLES = ("color", "point", "vector", "normal")
CODE2BASE = {"i": "int", "f": "float", "s": "string", "c": "color", "p": "point",
             "v": "vector", "n": "normal", "m": "matrix", "x": "void",
             "C": "closure color"}
BASE2CODE = {v: k for k, v in CODE2BASE.items()}


class T:
    __slots__ = ("base", "arr")

    def __init__(self, base, arr=0):
        self.base = base
        self.arr = arr  # 0 = not an array, -1 = unsized, n = length

    def __eq__(self, o):
        return isinstance(o, T) and self.base == o.base and self.arr == o.arr

    def __hash__(self):
        return hash((self.base, self.arr))

    def __repr__(self):
        if self.arr > 0:
            return "%s[%d]" % (self.base, self.arr)
        if self.arr < 0:
            return "%s[]" % self.base
        return self.base

    def elem(self):
        return T(self.base)

    @property
    def is_array(self):
        return self.arr != 0

    @property
    def is_triple(self):
        return self.base in TRIPLES and not self.arr

    @property
    def is_vectriple(self):
        return self.base in ("point", "vector", "normal") and not self.arr

    @property
    def is_float(self):
        return self.base == "float" and not self.arr

    @property
    def is_int(self):
        return self.base == "int" and not self.arr

    @property
    def is_string(self):
        return self.base == "string" and not self.arr

    @property
    def is_matrix(self):
        return self.base == "matrix" and not self.arr

    @property
    def is_closure(self):
        return self.base == "closure color"

    @property
    def is_void(self):
        return self.base == "void"

    @property
    def is_numeric(self):
        return not self.arr and self.base in ("int", "float", "matrix") + TRIPLES

    @property
    def is_int_or_float(self):
        return not self.arr and self.base in ("int", "float")

    @property
    def is_float_based(self):
        return self.base in ("float", "matrix") + TRIPLES

    def nfloats(self):
        n = {"int": 1, "float": 1, "string": 1, "matrix": 16}.get(self.base, 3)
        return n

    def aggregate(self):
        return {"matrix": 16}.get(self.base, 3 if self.base in TRIPLES else 1)


TINT, TFLOAT, TSTRING, TVOID = T("int"), T("float"), T("string"), T("void")
TCOLOR, TPOINT, TVECTOR, TNORMAL, TMATRIX = (T("color"), T("point"), T("vector"),
                                             T("normal"), T("matrix"))
TCLOSURE = T("closure color")


def equivalent(a, b):
    if a == b:
        return True
    if a.is_closure != b.is_closure:
        return False
    if a.base in TRIPLES and b.base in TRIPLES:
        return a.arr == b.arr or (a.arr < 0) != (b.arr < 0)
    return a.base == b.base and (a.arr == b.arr or ((a.arr < 0) != (b.arr < 0) and a.arr and b.arr))


def assignable(a, b):
    """can `a = b` ?  (src/liboslexec/typespec.cpp assignable())"""
    if equivalent(a, b):
        return True
    if a.is_closure:
        return b.is_closure or b.is_int  # closure = 0
    if b.is_closure:
        return False
    if a.arr or b.arr:
        # arrays of one element type assign whatever their lengths: the shorter length is copied
        # (testsuite/array-copy: int[3] = int[2])
        return a.arr > 0 and b.arr > 0 and (a.base == b.base or (a.base in TRIPLES and b.base in TRIPLES))
    if a.is_float:
        return b.is_int
    if a.is_triple:
        return b.is_int_or_float or b.is_triple
    if a.is_matrix:
        return b.is_int_or_float
    return False


class CompileError(Exception):
    pass


# ----------------------------------------------------------------------------
# preprocessor
# ----------------------------------------------------------------------------
def strip_comments(text):
    out = []
    i, n = 0, len(text)
    while i < n:
        c = text[i]
        if c == '"':
            j = i + 1
            while j < n and text[j] != '"':
                j += 2 if text[j] == "\\" else 1
            out.append(text[i:j + 1])
            i = j + 1
        elif text.startswith("//", i):
            j = text.find("\n", i)
            i = n if j < 0 else j
        elif text.startswith("/*", i):
            j = text.find("*/", i + 2)
            seg = text[i:j + 2]
            out.append("\n" * seg.count("\n") + " ")
            i = j + 2
        else:
            out.append(c)
            i += 1
    return "".join(out)


IDENT_RE = re.compile(r"[A-Za-z_]\w*")


class Preprocessor:
    def __init__(self, include_dirs, defines=None):
        self.include_dirs = include_dirs
        self.macros = dict(defines or {})  # name -> (params|None, body)
        # what oslc predefines (liboslcomp/oslcomp.cpp preprocess_buffer): the reference tree is 1.16.0
        for name, body in (("OSL_VERSION_MAJOR", "1"), ("OSL_VERSION_MINOR", "16"), ("OSL_VERSION_PATCH", "0"),
                           ("OSL_VERSION", "11600")):
            self.macros.setdefault(name, (None, body))
        self.included = set()

    def find(self, name, curdir):
        for d in [curdir] + self.include_dirs:
            p = os.path.join(d, name)
            if os.path.exists(p):
                return p
        raise CompileError("cannot find include file %s" % name)

    def expand(self, text, hide=frozenset()):
        out = []
        i, n = 0, len(text)
        while i < n:
            c = text[i]
            if c == '"':
                j = i + 1
                while j < n and text[j] != '"':
                    j += 2 if text[j] == "\\" else 1
                out.append(text[i:j + 1])
                i = j + 1
                continue
            m = IDENT_RE.match(text, i)
            if not m:
                # skip numbers whole so 1e5 / 0x.. aren't misparsed as idents
                m2 = re.match(r"\.?\d[\w.]*", text[i:])
                if m2 and (i == 0 or not (text[i - 1].isalnum() or text[i - 1] == "_")):
                    out.append(m2.group(0))
                    i += len(m2.group(0))
                else:
                    out.append(c)
                    i += 1
                continue
            name = m.group(0)
            i = m.end()
            if name in self.macros and name not in hide:
                params, body = self.macros[name]
                if params is None:
                    out.append(self.expand(body, hide | {name}))
                    continue
                # function-like: need '('
                j = i
                while j < n and text[j] in " \t\n":
                    j += 1
                if j < n and text[j] == "(":
                    depth, k, args, cur = 0, j, [], []
                    while k < n:
                        ch = text[k]
                        if ch == "(":
                            depth += 1
                            if depth > 1:
                                cur.append(ch)
                        elif ch == ")":
                            depth -= 1
                            if depth == 0:
                                break
                            cur.append(ch)
                        elif ch == "," and depth == 1:
                            args.append("".join(cur))
                            cur = []
                        else:
                            cur.append(ch)
                        k += 1
                    args.append("".join(cur))
                    if len(params) == 0 and args == [""]:
                        args = []
                    args = [self.expand(a.strip(), hide) for a in args]
                    sub = dict(zip(params, args))

                    def repl(mm):
                        return sub.get(mm.group(0), mm.group(0))
                    b = re.sub(r"#\s*([A-Za-z_]\w*)",
                               lambda mm: '"%s"' % sub.get(mm.group(1), mm.group(1)), body) \
                        if "#" in body.replace("##", "") else body
                    b = IDENT_RE.sub(repl, b)
                    b = re.sub(r"\s*##\s*", "", b)
                    out.append(self.expand(b, hide | {name}))
                    i = k + 1
                    continue
            out.append(name)
        return "".join(out)

    def eval_cond(self, expr):
        expr = re.sub(r"defined\s*\(\s*(\w+)\s*\)|defined\s+(\w+)",
                      lambda m: "1" if (m.group(1) or m.group(2)) in self.macros else "0", expr)
        expr = self.expand(expr)
        expr = IDENT_RE.sub("0", expr)
        expr = expr.replace("&&", " and ").replace("||", " or ")
        expr = re.sub(r"!(?!=)", " not ", expr)
        try:
            return bool(eval(expr, {"__builtins__": {}}, {}))
        except Exception:
            raise CompileError("cannot evaluate #if %s" % expr)

    def run(self, text, curdir="."):
        text = strip_comments(text).replace("\\\n", " ")
        lines = text.split("\n")
        out = []
        stack = []  # (active, taken_any, parent_active)
        active = True
        i = 0
        while i < len(lines):
            line = lines[i]
            i += 1
            s = line.strip()
            if s.startswith("#"):
                m = re.match(r"#\s*(\w+)\s*(.*)", s)
                if not m:
                    continue
                d, rest = m.group(1), m.group(2).strip()
                if d in ("ifdef", "ifndef", "if"):
                    if d == "ifdef":
                        c = rest.split()[0] in self.macros
                    elif d == "ifndef":
                        c = rest.split()[0] not in self.macros
                    else:
                        c = self.eval_cond(rest) if active else False
                    stack.append((active, c))
                    active = active and c
                elif d == "elif":
                    parent, taken = stack[-1]
                    c = (not taken) and parent and self.eval_cond(rest)
                    stack[-1] = (parent, taken or c)
                    active = parent and c
                elif d == "else":
                    parent, taken = stack[-1]
                    active = parent and not taken
                    stack[-1] = (parent, True)
                elif d == "endif":
                    active, _ = stack.pop()
                elif not active:
                    pass
                elif d == "define":
                    mm = re.match(r"(\w+)(\(([^)]*)\))?\s*(.*)", rest)
                    name = mm.group(1)
                    # function-like only if '(' immediately follows the name
                    if mm.group(2) is not None and rest[len(name):len(name) + 1] == "(":
                        params = [p.strip() for p in mm.group(3).split(",") if p.strip()]
                        self.macros[name] = (params, mm.group(4))
                    else:
                        self.macros[name] = (None, rest[len(name):].strip())
                elif d == "undef":
                    self.macros.pop(rest.split()[0], None)
                elif d == "include":
                    fn = rest.strip().strip('"<>')
                    p = self.find(fn, curdir)
                    if os.path.basename(p) == "stdosl.h" and "STDOSL_H" in self.macros:
                        continue
                    with open(p) as f:
                        out.append(self.run(f.read(), os.path.dirname(p)))
                # pragma/error/etc ignored
                continue
            if not active:
                continue
            # join lines while parentheses are unbalanced and a function-like
            # macro may be spanning lines
            while line.count("(") > line.count(")") and i < len(lines) \
                    and not lines[i].strip().startswith("#"):
                line += "\n" + lines[i]
                i += 1
            out.append(self.expand(line))
        return "\n".join(out)


# ----------------------------------------------------------------------------
# lexer
# ----------------------------------------------------------------------------
TOKEN_RE = re.compile(r"""
    (?P<float>(\d+\.\d*|\.\d+)([eE][-+]?\d+)?[fF]?|\d+[eE][-+]?\d+[fF]?)
  | (?P<hex>0[xX][0-9a-fA-F]+)
  | (?P<int>\d+)
  | (?P<id>[A-Za-z_]\w*)
  | (?P<str>"(\\.|[^"\\])*")
  | (?P<op>\[\[|\]\]|<<=|>>=|\+\+|--|&&|\|\||==|!=|<=|>=|<<|>>|\+=|-=|\*=|/=|&=|\|=|\^=|[-+*/%<>=!~&|^?:;,.(){}\[\]])
  | (?P<ws>\s+)
""", re.X)

KEYWORDS = {"if", "else", "for", "while", "do", "break", "continue", "return",
            "output", "closure", "struct", "public", "and", "or", "not"}
TYPENAMES = {"int", "float", "string", "color", "point", "vector", "normal",
             "matrix", "void"}
SHADERTYPES = {"shader", "surface", "displacement", "volume", "light"}


def unescape(s):
    return (s.replace("\\n", "\n").replace("\\t", "\t").replace('\\"', '"')
            .replace("\\\\", "\\"))


def lex(text):
    toks = []
    pos = 0
    while pos < len(text):
        m = TOKEN_RE.match(text, pos)
        if not m:
            raise CompileError("lex error near %r" % text[pos:pos + 30])
        pos = m.end()
        k = m.lastgroup
        v = m.group(k)
        if k == "ws":
            continue
        if k == "float":
            toks.append(("float", float(v.rstrip("fF"))))
        elif k == "hex":
            h = int(v, 16) & 0xffffffff        # OSL ints are 32 bit: 0xffffffff is -1
            toks.append(("int", h - (1 << 32) if h & 0x80000000 else h))
        elif k == "int":
            toks.append(("int", int(v)))
        elif k == "str":
            s = unescape(v[1:-1])
            # adjacent string literal concatenation
            if toks and toks[-1][0] == "str":
                toks[-1] = ("str", toks[-1][1] + s)
            else:
                toks.append(("str", s))
        elif k == "id":
            if v == "and":
                toks.append(("op", "&&"))
            elif v == "or":
                toks.append(("op", "||"))
            elif v == "not":
                toks.append(("op", "!"))
            else:
                toks.append(("id", v))
        else:
            toks.append(("op", v))
    toks.append(("eof", None))
    return toks


# ----------------------------------------------------------------------------
# AST
# ----------------------------------------------------------------------------
class N:
    def __init__(self, kind, **kw):
        self.kind = kind
        self.t = None
        self.__dict__.update(kw)


class Parser:
    def __init__(self, toks):
        self.toks = toks
        self.i = 0

    def peek(self, k=0):
        return self.toks[self.i + k]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def at(self, v):
        t = self.toks[self.i]
        return t[0] == "op" and t[1] == v

    def at_id(self, v=None):
        t = self.toks[self.i]
        return t[0] == "id" and (v is None or t[1] == v)

    def accept(self, v):
        if self.at(v):
            self.i += 1
            return True
        return False

    def expect(self, v):
        if not self.accept(v):
            raise CompileError("expected %r, got %r (tok %d)" % (v, self.peek(), self.i))

    def ident(self):
        t = self.next()
        if t[0] != "id":
            raise CompileError("expected identifier, got %r" % (t,))
        return t[1]

    # --- types -------------------------------------------------------------
    def at_type(self, k=0):
        t = self.peek(k)
        return t[0] == "id" and (t[1] in TYPENAMES or t[1] == "closure")

    def parse_type(self):
        if self.at_id("closure"):
            self.next()
            if self.ident() != "color":
                raise CompileError("only closure color supported")
            return T("closure color")
        name = self.ident()
        if name not in TYPENAMES:
            raise CompileError("unknown type %s (structs unsupported)" % name)
        return T(name)

    def parse_metadata(self):
        meta = []
        if self.accept("[["):
            while not self.at("]]"):
                t = self.parse_type()
                name = self.ident()
                self.expect("=")
                if self.at("{"):
                    val = self.parse_initlist()
                else:
                    val = self.parse_assign()
                meta.append((t, name, val))
                if not self.accept(","):
                    break
            self.expect("]]")
        return meta

    # --- top level ---------------------------------------------------------
    def parse_file(self):
        funcs, shader = [], None
        while self.peek()[0] != "eof":
            if self.at(";"):
                self.next()
                continue
            if self.at_id() and self.peek()[1] in SHADERTYPES and self.peek(1)[0] == "id":
                shader = self.parse_shader()
            elif self.at_id("struct"):
                raise CompileError("structs are not supported by mini_oslc")
            else:
                funcs.append(self.parse_function())
        return funcs, shader

    def parse_shader(self):
        stype = self.ident()
        name = self.ident()
        meta = self.parse_metadata()
        self.expect("(")
        params = self.parse_formals(shader=True)
        self.expect(")")
        body = self.parse_block()
        return N("shader", stype=stype, name=name, meta=meta, params=params, body=body)

    def parse_formals(self, shader=False):
        params = []
        while not self.at(")"):
            out = False
            if self.at_id("output"):
                self.next()
                out = True
            t = self.parse_type()
            name = self.ident()
            arr = 0
            if self.accept("["):
                if self.at("]"):
                    arr = -1
                else:
                    arr = self.next()[1]
                self.expect("]")
            init = None
            if self.accept("="):
                init = self.parse_initlist() if self.at("{") else self.parse_assign()
            meta = self.parse_metadata()
            params.append(N("param", type=T(t.base, arr), name=name, out=out,
                            init=init, meta=meta))
            if not self.accept(","):
                break
        return params

    def parse_function(self):
        rt = self.parse_type()
        name = self.ident()
        self.expect("(")
        formals = self.parse_formals()
        self.expect(")")
        meta = self.parse_metadata()
        body = None
        if not self.accept(";"):
            body = self.parse_block()
        return N("func", rtype=rt, name=name, formals=formals, meta=meta, body=body)

    # --- statements --------------------------------------------------------
    def parse_block(self):
        self.expect("{")
        stmts = []
        while not self.at("}"):
            stmts.append(self.parse_stmt())
        self.expect("}")
        return N("block", stmts=stmts)

    def is_funcdecl(self):
        # type ident ( ... ) {   inside a function body
        if not self.at_type():
            return False
        k = 2 if self.peek()[1] == "closure" else 1
        if self.peek(k)[0] != "id" or self.peek(k + 1) != ("op", "("):
            return False
        depth, j = 0, self.i + k + 1
        while True:
            t = self.toks[j]
            if t == ("op", "("):
                depth += 1
            elif t == ("op", ")"):
                depth -= 1
                if depth == 0:
                    break
            elif t[0] == "eof":
                return False
            j += 1
        nt = self.toks[j + 1]
        return nt == ("op", "{") or nt == ("op", "[[")

    def parse_stmt(self):
        if self.at("{"):
            return self.parse_block()
        if self.at(";"):
            self.next()
            return N("block", stmts=[])
        if self.at_id("if"):
            self.next()
            self.expect("(")
            c = self.parse_expr()
            self.expect(")")
            a = self.parse_stmt()
            b = None
            if self.at_id("else"):
                self.next()
                b = self.parse_stmt()
            return N("if", cond=c, a=a, b=b)
        if self.at_id("for"):
            self.next()
            self.expect("(")
            init = None
            if not self.at(";"):
                init = self.parse_vardecl() if self.at_type() else N("expr", e=self.parse_expr())
                if init.kind == "expr":
                    self.expect(";")
            else:
                self.next()
            cond = None if self.at(";") else self.parse_expr()
            self.expect(";")
            it = None if self.at(")") else self.parse_expr()
            self.expect(")")
            body = self.parse_stmt()
            return N("loop", op="for", init=init, cond=cond, iter=it, body=body)
        if self.at_id("while"):
            self.next()
            self.expect("(")
            c = self.parse_expr()
            self.expect(")")
            return N("loop", op="while", init=None, cond=c, iter=None, body=self.parse_stmt())
        if self.at_id("do"):
            self.next()
            body = self.parse_stmt()
            if not self.at_id("while"):
                raise CompileError("expected while")
            self.next()
            self.expect("(")
            c = self.parse_expr()
            self.expect(")")
            self.expect(";")
            return N("loop", op="dowhile", init=None, cond=c, iter=None, body=body)
        if self.at_id("break") or self.at_id("continue"):
            k = self.next()[1]
            self.expect(";")
            return N("loopmod", op=k)
        if self.at_id("return"):
            self.next()
            e = None if self.at(";") else self.parse_expr()
            self.expect(";")
            return N("return", e=e)
        if self.is_funcdecl():
            return N("funcdecl", f=self.parse_function())
        if self.at_type() and not self.peek(1) == ("op", "("):
            return self.parse_vardecl()
        e = self.parse_expr()
        self.expect(";")
        return N("expr", e=e)

    def parse_vardecl(self):
        t = self.parse_type()
        decls = []
        while True:
            name = self.ident()
            arr = 0
            if self.accept("["):
                if self.at("]"):
                    arr = -1
                else:
                    arr = self.next()[1]
                self.expect("]")
            init = None
            if self.accept("="):
                init = self.parse_initlist() if self.at("{") else self.parse_assign()
            decls.append(N("vardecl", type=T(t.base, arr), name=name, init=init))
            if not self.accept(","):
                break
        self.expect(";")
        return decls[0] if len(decls) == 1 else N("block", stmts=decls, noscope=True)

    def parse_initlist(self):
        self.expect("{")
        items = []
        while not self.at("}"):
            items.append(self.parse_initlist() if self.at("{") else self.parse_assign())
            if not self.accept(","):
                break
        self.expect("}")
        return N("initlist", items=items)

    # --- expressions -------------------------------------------------------
    def parse_expr(self):
        e = self.parse_assign()
        while self.accept(","):
            r = self.parse_assign()
            e = N("comma", l=e, r=r)
        return e

    ASSIGN_OPS = {"=": None, "+=": "add", "-=": "sub", "*=": "mul", "/=": "div",
                  "&=": "bitand", "|=": "bitor", "^=": "xor", "<<=": "shl", ">>=": "shr"}

    def parse_assign(self):
        lhs = self.parse_ternary()
        t = self.peek()
        if t[0] == "op" and t[1] in self.ASSIGN_OPS:
            self.next()
            rhs = self.parse_initlist() if self.at("{") else self.parse_assign()
            return N("assign", op=self.ASSIGN_OPS[t[1]], lhs=lhs, rhs=rhs)
        return lhs

    def parse_ternary(self):
        c = self.parse_binary(0)
        if self.accept("?"):
            a = self.parse_expr()
            self.expect(":")
            b = self.parse_assign()
            return N("ternary", cond=c, a=a, b=b)
        return c

    BINOPS = [
        {"||": "or"}, {"&&": "and"}, {"|": "bitor"}, {"^": "xor"}, {"&": "bitand"},
        {"==": "eq", "!=": "neq"}, {"<": "lt", ">": "gt", "<=": "le", ">=": "ge"},
        {"<<": "shl", ">>": "shr"}, {"+": "add", "-": "sub"},
        {"*": "mul", "/": "div", "%": "mod"},
    ]

    def parse_binary(self, lvl):
        if lvl >= len(self.BINOPS):
            return self.parse_unary()
        l = self.parse_binary(lvl + 1)
        while True:
            t = self.peek()
            if t[0] == "op" and t[1] in self.BINOPS[lvl]:
                self.next()
                r = self.parse_binary(lvl + 1)
                l = N("binary", op=self.BINOPS[lvl][t[1]], l=l, r=r)
            else:
                return l

    def parse_unary(self):
        t = self.peek()
        if t[0] == "op":
            if t[1] in ("-", "+", "!", "~"):
                self.next()
                e = self.parse_unary()
                op = {"-": "neg", "+": "pos", "!": "not", "~": "compl"}[t[1]]
                if op == "neg" and e.kind == "lit" and e.t.base in ("int", "float"):
                    return N("lit", t=e.t, value=-e.value)
                return N("unary", op=op, e=e)
            if t[1] in ("++", "--"):
                self.next()
                e = self.parse_unary()
                return N("preinc", op="add" if t[1] == "++" else "sub", e=e)
            if t[1] == "(" and self.at_type(1):
                # typecast:  ( type ) expr
                k = 3 if self.peek(1)[1] == "closure" else 2
                if self.peek(k) == ("op", ")"):
                    self.next()
                    ty = self.parse_type()
                    self.expect(")")
                    return N("cast", type=ty, e=self.parse_unary())
        return self.parse_postfix()

    def parse_postfix(self):
        e = self.parse_primary()
        while True:
            if self.accept("["):
                idx = self.parse_expr()
                self.expect("]")
                if e.kind == "index" and len(e.idx) < 3:
                    e.idx.append(idx)
                else:
                    e = N("index", base=e, idx=[idx])
            elif self.at(".") and self.peek(1)[0] == "id":
                self.next()
                f = self.ident()
                comp = {"x": 0, "y": 1, "z": 2, "r": 0, "g": 1, "b": 2}.get(f)
                if comp is None:
                    raise CompileError("struct field .%s unsupported" % f)
                lit = N("lit", t=TINT, value=comp)
                if e.kind == "index" and len(e.idx) < 3:
                    e.idx.append(lit)
                else:
                    e = N("index", base=e, idx=[lit])
            elif self.at("++") or self.at("--"):
                op = self.next()[1]
                e = N("postinc", op="add" if op == "++" else "sub", e=e)
            else:
                return e

    def parse_primary(self):
        t = self.next()
        if t[0] == "int":
            return N("lit", t=TINT, value=t[1])
        if t[0] == "float":
            return N("lit", t=TFLOAT, value=t[1])
        if t[0] == "str":
            return N("lit", t=TSTRING, value=t[1])
        if t == ("op", "("):
            e = self.parse_expr()
            self.expect(")")
            return e
        if t[0] == "id":
            name = t[1]
            if self.at("("):
                self.next()
                args = []
                while not self.at(")"):
                    args.append(self.parse_initlist() if self.at("{") else self.parse_assign())
                    if not self.accept(","):
                        break
                self.expect(")")
                if name in TYPENAMES:
                    return N("ctor", type=T(name), args=args)
                return N("call", name=name, args=args)
            return N("var", name=name)
        raise CompileError("unexpected token %r at %d" % (t, self.i))


# ----------------------------------------------------------------------------
# symbols / IR
# ----------------------------------------------------------------------------
class Sym:
    def __init__(self, name, t, symtype, value=None):
        self.name = name       # mangled name as written to the .oso
        self.t = t
        self.symtype = symtype  # param oparam local temp global const
        self.value = value      # consts / param defaults: list of python values
        self.alias = None
        self.initexpr = False
        self.meta = []
        self.used = False

    def deref(self):
        s = self
        while s.alias is not None:
            s = s.alias
        return s


class Op:
    def __init__(self, name, args, rw, jumps=(), method="___main___", derivs=()):
        self.name = name
        self.args = args
        self.rw = rw
        self.jumps = list(jumps)
        self.method = method
        self.derivs = derivs


GLOBALS = {"P": TPOINT, "I": TVECTOR, "N": TNORMAL, "Ng": TNORMAL, "u": TFLOAT,
           "v": TFLOAT, "dPdu": TVECTOR, "dPdv": TVECTOR, "Ps": TPOINT,
           "time": TFLOAT, "dtime": TFLOAT, "dPdtime": TVECTOR, "Ci": TCLOSURE}

# argcode tables for the builtins that are NOT declared in stdosl.h
# (the language's intrinsic table; typecheck.cpp:2130-2186)
NOISE_ARGS = ["ff", "fff", "fp", "fpf", "cf", "cff", "cp", "cpf", "vf", "vff", "vp", "vpf"]
PNOISE_ARGS = ["fff", "fffff", "fpp", "fpfpf", "cff", "cffff", "cpp", "cpfpf",
               "vff", "vffff", "vpp", "vpfpf"]
GNOISE_ARGS = ["fsf.", "fsff.", "fsp.", "fspf.", "csf.", "csff.", "csp.", "cspf.",
               "vsf.", "vsff.", "vsp.", "vspf."]
PGNOISE_ARGS = ["fsff.", "fsffff.", "fspp.", "fspfpf.", "csff.", "csffff.", "cspp.",
                "cspfpf.", "vsff.", "vsffff.", "vspp.", "vspfpf."]
INTRINSICS = {
    "area": ["fp"], "arraylength": ["i?[]"], "calculatenormal": ["vp"],
    "cellnoise": NOISE_ARGS, "hashnoise": NOISE_ARGS, "snoise": NOISE_ARGS,
    "noise": GNOISE_ARGS + NOISE_ARGS, "pnoise": PGNOISE_ARGS + PNOISE_ARGS,
    "psnoise": PNOISE_ARGS, "concat": ["sss"],
    "Dx": ["ff", "vp", "vv", "vn", "cc"], "Dy": ["ff", "vp", "vv", "vn", "cc"],
    "Dz": ["ff", "vp", "vv", "vn", "cc"], "filterwidth": ["ff", "vp", "vv"],
    "error": ["xs*"], "warning": ["xs*"], "printf": ["xs*"], "format": ["ss*"],
    "fprintf": ["xss*"], "exit": ["x"],
    "getattribute": ["is?", "is?[]", "iss?", "iss?[]", "isi?", "isi?[]", "issi?", "issi?[]"],
    "getmessage": ["is?", "is?[]", "iss?", "iss?[]"], "setmessage": ["xs?", "xs?[]"],
    "isconnected": ["i?"], "isconstant": ["i?"],
    "random": ["f", "c", "p", "v", "n"],
    "sincos": ["xfff", "xccc", "xppp", "xvvv", "xnnn"],
    "spline": ["fsff[]", "csfc[]", "psfp[]", "vsfv[]", "nsfn[]", "fsfif[]", "csfic[]",
               "psfip[]", "vsfiv[]", "nsfin[]"],
    "splineinverse": ["fsff[]", "fsfif[]"], "surfacearea": ["f"],
    "texture": ["fsff.", "fsffffff.", "csff.", "csffffff.", "vsff.", "vsffffff."],
    "environment": ["fsv.", "fsvvv.", "csv.", "csvvv.", "vsv.", "vsvvv."],
    "trace": ["ipv."], "bump": ["xf", "xsf", "xv"], "displace": ["xf", "xsf", "xv"],
    "gettextureinfo": ["iss?", "iss?[]", "isffs?", "isffs?[]"],
}
TAKES_DERIVS = {"area", "calculatenormal", "Dx", "Dy", "Dz", "filterwidth", "texture",
                "environment", "trace", "noise", "pnoise", "bump", "displace"}


def parse_argcodes(code):
    """'fpf' -> (rettype, [argspec...]) ; argspec = T | '*' | '.' | '?' | '?[]'"""
    out = []
    i = 0
    while i < len(code):
        c = code[i]
        if c in "*.":
            out.append(c)
            i += 1
        elif c == "?":
            if code[i + 1:i + 3] == "[]":
                out.append("?[]")
                i += 3
            else:
                out.append("?")
                i += 1
        else:
            t = T(CODE2BASE[c])
            i += 1
            if code[i:i + 1] == "[":
                j = code.index("]", i)
                t = T(t.base, int(code[i + 1:j]) if j > i + 1 else -1)
                i = j + 1
            out.append(t)
    return out[0], out[1:]


def type_code(t):
    c = BASE2CODE[t.base]
    if t.arr > 0:
        c += "[%d]" % t.arr
    elif t.arr < 0:
        c += "[]"
    return c


class Func:
    def __init__(self, name, rtype, formals, builtin, node=None):
        self.name = name
        self.rtype = rtype
        self.formals = formals   # list of argspecs (T or wildcard strings)
        self.builtin = builtin
        self.node = node         # AST for user functions
        self.argcodes = type_code(rtype) + "".join(
            f if isinstance(f, str) else type_code(f) for f in formals)


def fmt_float(f):
    s = "%.9g" % f
    return s


# ----------------------------------------------------------------------------
# compiler
# ----------------------------------------------------------------------------
class Compiler:
    def __init__(self):
        self.funcs = {}          # name -> [Func]
        self.scopes = [{}]
        self.syms = []           # ordered symbols
        self.ops = []
        self.consts = {}
        self.ntemps = 0
        self.method = "___main___"
        self.func_stack = []     # (Func, return_sym)
        self.scope_id = 0
        self.names = set()
        for name, codes in INTRINSICS.items():
            for c in codes:
                rt, fm = parse_argcodes(c)
                self.add_func(Func(name, rt, fm, True))

    # -- symbol helpers -----------------------------------------------------
    def add_func(self, f):
        lst = self.funcs.setdefault(f.name, [])
        for g in lst:
            if g.argcodes == f.argcodes:
                if f.node is not None and f.node.body is not None:
                    lst[lst.index(g)] = f
                return
        lst.append(f)

    def lookup(self, name):
        for sc in reversed(self.scopes):
            if name in sc:
                return sc[name]
        if name in GLOBALS:
            s = Sym(name, GLOBALS[name], "global")
            self.scopes[0][name] = s
            self.syms.append(s)
            return s
        raise CompileError("unknown identifier '%s'" % name)

    def declare(self, name, t, symtype):
        mangled = name
        if symtype == "local":
            if name in self.names or name in GLOBALS:
                self.scope_id += 1
                mangled = "___%d_%s" % (self.scope_id + 300, name)
        self.names.add(mangled)
        s = Sym(mangled, t, symtype)
        self.scopes[-1][name] = s
        self.syms.append(s)
        return s

    def temp(self, t):
        self.ntemps += 1
        s = Sym("$tmp%d" % self.ntemps, t, "temp")
        self.syms.append(s)
        return s

    def const(self, t, vals):
        if not isinstance(vals, (list, tuple)):
            vals = [vals]
        if t.base != "string" and t.base != "int":
            vals = [struct.unpack("f", struct.pack("f", float(v)))[0] for v in vals]
        key = (repr(t), tuple(repr(v) for v in vals))
        if key in self.consts:
            return self.consts[key]
        s = Sym("$const%d" % (len(self.consts) + 1), t, "const", list(vals))
        self.consts[key] = s
        self.syms.append(s)
        return s

    def emit(self, name, args, rw=None, jumps=(), derivs=()):
        args = [a.deref() for a in args]
        if rw is None:
            rw = "w" + "r" * (len(args) - 1) if args else ""
        self.ops.append(Op(name, args, rw, jumps, self.method, derivs))
        return len(self.ops) - 1

    def next_label(self):
        return len(self.ops)

    # -- typecheck ----------------------------------------------------------
    def tc(self, n, expected=None):
        m = getattr(self, "tc_" + n.kind)
        n.t = m(n, expected)
        return n.t

    def tc_lit(self, n, e):
        return n.t

    def tc_var(self, n, e):
        n.sym = self.lookup(n.name)
        return n.sym.t

    def tc_comma(self, n, e):
        self.tc(n.l, e)
        return self.tc(n.r, e)

    def tc_unary(self, n, e):
        t = self.tc(n.e, e)
        if n.op == "not":
            return TINT
        return t

    def tc_preinc(self, n, e):
        return self.tc(n.e)

    tc_postinc = tc_preinc

    def tc_binary(self, n, e):
        l = self.tc(n.l, e)
        r = self.tc(n.r, e)
        op = n.op
        if l.is_closure or r.is_closure:
            if op == "add" and l.is_closure and r.is_closure:
                return l
            if op == "mul" and l.is_closure != r.is_closure:
                if r.is_closure:
                    n.l, n.r = n.r, n.l
                return TCLOSURE
            if op in ("and", "or"):
                return TINT
            raise CompileError("bad closure op %s" % op)
        if op in ("add", "sub", "mul", "div"):
            if equivalent(l, r):
                if op == "sub" and l.base == "point" and r.base == "point":
                    return TVECTOR
                if op in ("add", "sub") and (l.base == "point" or r.base == "point"):
                    return TPOINT
                return l
            if (l.is_numeric and r.is_int_or_float) or (l.is_int_or_float and r.is_numeric):
                if l.aggregate() > r.aggregate():
                    return l
                if r.aggregate() > l.aggregate():
                    return r
                return r if r.base == "float" else l
        elif op == "mod":
            if l.is_int and r.is_int:
                return TINT
        elif op in ("eq", "neq"):
            if equivalent(l, r) or (l.is_numeric and r.is_int_or_float) or \
                    (l.is_int_or_float and r.is_numeric):
                return TINT
        elif op in ("lt", "gt", "le", "ge"):
            if l.is_int_or_float and r.is_int_or_float:
                return TINT
        elif op in ("bitand", "bitor", "xor", "shl", "shr"):
            if l.is_int and r.is_int:
                return TINT
        elif op in ("and", "or"):
            return TINT
        raise CompileError("Not allowed: '%s %s %s'" % (l, op, r))

    def tc_ternary(self, n, e):
        self.tc(n.cond)
        t = self.tc(n.a, e)
        f = self.tc(n.b, e)
        if assignable(t, f):
            return t
        if assignable(f, t):
            return f
        raise CompileError("ternary type mismatch %s vs %s" % (t, f))

    def tc_cast(self, n, e):
        self.tc(n.e, n.type)
        return n.type

    def tc_assign(self, n, e):
        lt = self.tc(n.lhs)
        if n.rhs.kind == "initlist":
            self.tc_initlist_as(n.rhs, lt)
        else:
            rt = self.tc(n.rhs, lt)
            if n.op is None and not assignable(lt, rt):
                raise CompileError("cannot assign %s = %s" % (lt, rt))
        return lt

    def tc_initlist_as(self, n, t):
        n.t = t
        if t.is_array:
            for it in n.items:
                if it.kind == "initlist":
                    self.tc_initlist_as(it, t.elem())
                else:
                    self.tc(it, t.elem())
        else:
            for it in n.items:
                self.tc(it, TFLOAT if t.is_float_based else t)
        return t

    def tc_initlist(self, n, e):
        if e is None:
            raise CompileError("initializer list without a type")
        return self.tc_initlist_as(n, e)

    def tc_index(self, n, e):
        bt = self.tc(n.base)
        for ix in n.idx:
            self.tc(ix)
        k = len(n.idx)
        t = bt
        if t.is_array:
            t = t.elem()
            k -= 1
        if k == 0:
            return t
        if t.is_triple and k == 1:
            return TFLOAT
        if t.is_matrix and k == 2:
            return TFLOAT
        raise CompileError("bad indexing of %s" % bt)

    def tc_ctor(self, n, e):
        t = n.type
        argexp = None
        if t.is_float:
            argexp = TFLOAT
        elif t.is_triple:
            argexp = e if (len(n.args) == 1 and e is not None and e.is_triple) else TFLOAT
            if len(n.args) == 1:
                argexp = t
        elif t.is_matrix:
            argexp = TFLOAT
        elif t.is_int:
            argexp = TINT
        for i, a in enumerate(n.args):
            if a.kind == "lit" and a.t.is_string:
                self.tc(a)
            else:
                self.tc(a, argexp)
        return t

    def score_type(self, exp, act):
        if exp == act:
            return 100
        if (not act.is_closure and act.is_int_or_float and not exp.is_closure
                and exp.is_int_or_float):
            return 0 if exp.is_int else 77
        if exp.arr < 0 and act.arr > 0 and exp.base == act.base:
            return 44
        if assignable(exp, act):
            if act.is_vectriple and exp.is_vectriple:
                return 32
            if act.is_triple and exp.is_triple:
                return 27
            return 23
        return 0

    def score_func(self, f, argtypes):
        score, i, n = 0, 0, len(argtypes)
        fi = 0
        formals = f.formals
        while fi < len(formals) and i < n:
            fm = formals[fi]
            if fm == "*":
                score += n - i
                i = n
                fi += 1
                continue
            if fm == ".":
                if argtypes[i].is_string and i + 1 < n:
                    score += n - i
                    i = n
                    fi += 1
                    continue
                return 0
            if fm == "?[]":
                if not argtypes[i].is_array:
                    return 0
                score += 1
            elif fm == "?":
                if argtypes[i].is_array:
                    return 0
                score += 1
            else:
                s = self.score_type(fm, argtypes[i])
                if s == 0:
                    return 0
                score += s
            i += 1
            fi += 1
        if fi < len(formals) and formals[fi] in ("*", "."):
            fi += 1
        if fi == len(formals) and i < n and f.rtype.is_closure and getattr(f, "builtin", True):
            # closure constructors take trailing "keyword", value pairs (oslc:
            # ASTfunction_call::check_arglist accepts them for closure-returning builtins)
            rest = argtypes[i:]
            if len(rest) % 2 == 0 and all(rest[k].is_string for k in range(0, len(rest), 2)):
                i = n
        if fi < len(formals) or i < n:
            return 0
        return max(score, 1)

    RANK = {"float": 0, "int": 1, "color": 2, "vector": 3, "point": 4, "normal": 5,
            "matrix": 6, "string": 7, "closure color": 8, "void": 10}

    def tc_call(self, n, e):
        if n.name not in self.funcs:
            raise CompileError("unknown function '%s'" % n.name)
        for a in n.args:
            if a.kind != "initlist":
                self.tc(a, e)
        argtypes = [a.t if a.kind != "initlist" else None for a in n.args]
        if any(t is None for t in argtypes):
            raise CompileError("initializer-list args unsupported")
        best, cands = 0, []
        for f in self.funcs[n.name]:
            s = self.score_func(f, argtypes)
            if s == 0 or s < best:
                continue
            if s > best:
                cands = []
                best = s
            cands.append(f)
        if not cands:
            raise CompileError("No matching function call to '%s(%s)'" % (
                n.name, ", ".join(map(repr, argtypes))))
        if len(cands) > 1:
            ev = e if e is not None else T("unknown")
            rs = [self.score_type(ev, f.rtype) if e is not None else 0 for f in cands]
            top = max(rs)
            tops = [f for f, r in zip(cands, rs) if r == top]
            if len(tops) > 1:
                tops.sort(key=lambda f: self.RANK[f.rtype.base])
            cands = tops
        n.func = cands[0]
        # re-typecheck args with the chosen formal types as the expected type so
        # that nested polymorphic calls resolve the way the formals want
        for a, fm in zip(n.args, n.func.formals):
            if isinstance(fm, T) and a.kind == "call":
                self.tc(a, fm)
        return n.func.rtype

    # -- codegen ------------------------------------------------------------
    def coerce(self, sym, t, acceptfloat=False):
        st = sym.t
        if equivalent(st, t) or t.arr < 0:
            return sym
        if sym.symtype == "const" and st.is_int and t.is_float_based and not t.is_array:
            if t.is_float or acceptfloat:
                return self.const(TFLOAT, float(sym.value[0]))
        if acceptfloat and st.is_float and t.is_float_based:
            return sym
        tmp = self.temp(t)
        self.emit("assign", [tmp, sym])
        return tmp

    def cg(self, n, dest=None):
        return getattr(self, "cg_" + n.kind)(n, dest)

    def cg_lit(self, n, dest):
        return self.const(n.t, n.value)

    def cg_var(self, n, dest):
        return n.sym.deref()

    def cg_comma(self, n, dest):
        self.cg(n.l)
        return self.cg(n.r, dest)

    def cg_int(self, n, boolify=False, invert=False):
        d = self.cg(n)
        if not d.t.is_int or boolify or invert:
            tmp = self.temp(TINT)
            if d.t.is_string:
                z = self.const(TSTRING, "")
            elif d.t.is_int or d.t.is_closure:
                z = self.const(TINT, 0)
            else:
                z = self.const(TFLOAT, 0.0)
            self.emit("eq" if invert else "neq", [tmp, d, z])
            d = tmp
        return d

    def cg_unary(self, n, dest):
        if n.op == "not":
            return self.cg_int(n.e, True, True)
        e = self.cg(n.e)
        if n.op == "pos":
            return e
        if dest is None or not equivalent(dest.t, n.t):
            dest = self.temp(n.t)
        if e.t.is_closure:
            self.emit("mul", [dest, e, self.const(TFLOAT, -1.0)])
            return dest
        self.emit(n.op, [dest, e])
        return dest

    def cg_incdec(self, n, dest, post):
        sym = self.cg(n.e)
        one = self.const(TINT, 1) if sym.t.is_int else self.const(TFLOAT, 1.0)
        old = None
        if post:
            old = dest if dest is not None and equivalent(dest.t, sym.t) else self.temp(sym.t)
            self.emit("assign", [old, sym])
        if n.e.kind == "index":
            tmp = self.temp(sym.t)
            self.emit(n.op, [tmp, sym, one])
            self.store_index(n.e, tmp)
            return old if post else tmp
        self.emit(n.op, [sym, sym, one])
        return old if post else sym

    def cg_preinc(self, n, dest):
        return self.cg_incdec(n, dest, False)

    def cg_postinc(self, n, dest):
        return self.cg_incdec(n, dest, True)

    def cg_binary(self, n, dest):
        if n.op in ("and", "or"):
            return self.cg_logic(n)
        l = self.cg(n.l)
        r = self.cg(n.r)
        if dest is None or not equivalent(dest.t, n.t):
            dest = self.temp(n.t)
        if n.t.is_closure:
            if n.op in ("mul", "div"):
                r = self.coerce(r, TCOLOR, True)
            self.emit(n.op, [dest, l, r])
            return dest
        if n.op in ("mul", "div", "add", "sub"):
            if l.t.is_float_based and r.t.is_int:
                if r.symtype == "const":
                    r = self.const(TFLOAT, float(r.value[0]))
                else:
                    tmp = self.temp(l.t)
                    self.emit("assign", [tmp, r])
                    r = tmp
            elif l.t.is_int and r.t.is_float_based:
                if l.symtype == "const":
                    l = self.const(TFLOAT, float(l.value[0]))
                else:
                    tmp = self.temp(r.t)
                    self.emit("assign", [tmp, l])
                    l = tmp
        self.emit(n.op, [dest, l, r])
        return dest

    def cg_logic(self, n):
        dest = self.cg_int(n.l, True)
        ifop = self.emit("if", [dest], "r")
        if n.op == "and":
            r = self.cg_int(n.r, True)
            if r is not dest:
                self.emit("assign", [dest, r])
            fl = self.next_label()
        else:
            fl = self.next_label()
            r = self.cg_int(n.r, True)
            if r is not dest:
                self.emit("assign", [dest, r])
        self.ops[ifop].jumps = [fl, self.next_label()]
        return dest

    def cg_ternary(self, n, dest):
        if dest is None or not equivalent(dest.t, n.t):
            dest = self.temp(n.t)
        c = self.cg_int(n.cond)
        ifop = self.emit("if", [c], "r")
        a = self.cg(n.a, dest)
        if a is not dest:
            self.emit("assign", [dest, a])
        fl = self.next_label()
        b = self.cg(n.b, dest)
        if b is not dest:
            self.emit("assign", [dest, b])
        self.ops[ifop].jumps = [fl, self.next_label()]
        return dest

    def cg_cast(self, n, dest):
        e = self.cg(n.e, dest)
        if equivalent(n.t, e.t):
            return e
        if dest is None or not equivalent(dest.t, n.t):
            dest = self.temp(n.t)
        self.emit("assign", [dest, e])
        return dest

    def cg_ctor(self, n, dest):
        t = n.t
        if t.is_triple and len(n.args) in (1, 3) and all(
                a.kind == "lit" and a.t.base in ("int", "float") for a in n.args):
            f = [float(a.value) for a in n.args]
            if len(f) == 1:
                f = f * 3
            return self.const(t, f)
        if dest is None or not equivalent(dest.t, t):
            dest = self.temp(t)
        argeval = dest if (t.is_float and len(n.args) == 1 and n.args[0].t.is_float) else None
        args = [dest]
        for a in n.args:
            v = self.cg(a, argeval)
            if v.t.is_int and not t.is_int:
                if v.symtype == "const":
                    v = self.const(TFLOAT, float(v.value[0]))
                else:
                    tmp = self.temp(TFLOAT)
                    self.emit("assign", [tmp, v])
                    v = tmp
            args.append(v)
        if len(n.args) == 1 and args[1] is dest:
            pass
        elif len(n.args) == 1:
            self.emit("assign", args)
        else:
            self.emit(t.base, args)
        return dest

    def cg_index(self, n, dest):
        base = self.cg(n.base)
        idx = [self.cg(ix) for ix in n.idx]
        if dest is None or not equivalent(dest.t, n.t):
            dest = self.temp(n.t)
        cur = base
        k = 0
        if cur.t.is_array:
            if len(idx) == 1:
                self.emit("aref", [dest, cur, idx[0]])
                return dest
            el = self.temp(cur.t.elem())
            self.emit("aref", [el, cur, idx[0]])
            cur = el
            k = 1
        if cur.t.is_triple:
            self.emit("compref", [dest, cur, idx[k]])
        elif cur.t.is_matrix:
            self.emit("mxcompref", [dest, cur, idx[k], idx[k + 1]])
        else:
            raise CompileError("cannot index %s" % cur.t)
        return dest

    def store_index(self, n, src):
        """n is an index node used as an lvalue; store src into it."""
        base = self.cg(n.base)
        idx = [self.cg(ix) for ix in n.idx]
        if base.t.is_array:
            if len(idx) == 1:
                src = self.coerce(src, base.t.elem(), True)
                self.emit("aassign", [base, idx[0], src], "wrr")
                return
            el = self.temp(base.t.elem())
            self.emit("aref", [el, base, idx[0]])
            self.store_comp(el, idx[1:], src)
            self.emit("aassign", [base, idx[0], el], "wrr")
            return
        self.store_comp(base, idx, src)

    def store_comp(self, base, idx, src):
        if src.t.is_int:
            src = self.coerce(src, TFLOAT)
        if base.t.is_triple:
            self.emit("compassign", [base, idx[0], src], "wrr")
        elif base.t.is_matrix:
            self.emit("mxcompassign", [base, idx[0], idx[1], src], "wrrr")
        else:
            raise CompileError("cannot index-assign %s" % base.t)

    def cg_assign(self, n, dest):
        lhs = n.lhs
        if lhs.kind == "index":
            if n.op is None:
                src = self.cg(n.rhs)
            else:
                cur = self.cg(lhs)
                r = self.cg(n.rhs)
                if cur.t.is_float_based and r.t.is_int:
                    r = self.coerce(r, TFLOAT)
                src = self.temp(lhs.t)
                self.emit(n.op, [src, cur, r])
            self.store_index(lhs, src)
            return src
        target = self.cg(lhs)
        if n.rhs.kind == "initlist":
            self.cg_initlist_into(n.rhs, target)
            return target
        if n.op is None:
            r = self.cg(n.rhs, target)
            if r is not target:
                if r.symtype == "const" and r.t.is_int and target.t.is_float_based \
                        and not target.t.is_array:
                    r = self.const(TFLOAT, float(r.value[0]))
                self.emit("assign", [target, r])
            return target
        r = self.cg(n.rhs)
        if target.t.is_float_based and r.t.is_int:
            if r.symtype == "const":
                r = self.const(TFLOAT, float(r.value[0]))
            else:
                tmp = self.temp(TFLOAT)
                self.emit("assign", [tmp, r])
                r = tmp
        if target.t.is_closure and n.op in ("mul", "div"):
            r = self.coerce(r, TCOLOR, True)
        self.emit(n.op, [target, target, r])
        return target

    def cg_initlist_into(self, n, target):
        t = target.t
        if t.is_array:
            for i, it in enumerate(n.items):
                if it.kind == "initlist":
                    tmp = self.temp(t.elem())
                    self.cg_initlist_into(it, tmp)
                    v = tmp
                else:
                    v = self.coerce(self.cg(it), t.elem(), True)
                self.emit("aassign", [target, self.const(TINT, i), v], "wrr")
        else:
            fake = N("ctor", type=t, args=n.items)
            fake.t = t
            r = self.cg_ctor(fake, target)
            if r is not target:
                self.emit("assign", [target, r])

    def cg_initlist(self, n, dest):
        if dest is None:
            dest = self.temp(n.t)
        self.cg_initlist_into(n, dest)
        return dest

    RW_SPECIAL = {"sincos": {1: "w", 2: "w"}}

    def cg_call(self, n, dest):
        f = n.func
        rt = f.rtype
        if not rt.is_void:
            if dest is None or not equivalent(dest.t, rt):
                dest = self.temp(rt)
        else:
            dest = None
        if not f.builtin:
            return self.cg_usercall(n, f, dest)
        name = n.name
        if name == "transform":
            if rt.base == "vector":
                name = "transformv"
            elif rt.base == "normal":
                name = "transformn"
        argsyms = []
        outfix = []
        for i, a in enumerate(n.args):
            fm = f.formals[i] if i < len(f.formals) else None
            s = self.cg(a)
            if isinstance(fm, T):
                s = self.coerce(s, fm)
            argsyms.append(s)
        args = list(argsyms)
        off = 0
        if rt.is_closure:
            args.insert(0, self.const(TSTRING, n.name))
            off += 1
        if dest is not None:
            args.insert(0, dest)
            off += 1
        rw = ["r"] * len(args)
        if dest is not None:
            rw[0] = "w"
        nargs = len(n.args)
        writes = []
        if name == "sincos":
            writes = [1, 2]
        elif name in ("getattribute", "getmessage", "gettextureinfo", "dict_value"):
            writes = [nargs - 1]
        for w in writes:
            rw[w + off] = "w"
        derivs = []
        if name in TAKES_DERIVS:
            if name in ("area", "filterwidth", "calculatenormal", "Dx", "Dy", "Dz"):
                derivs = [1]
            elif name == "texture":
                if nargs == 3 or n.args[3].t.is_string:
                    derivs = [2, 3]
            elif name == "environment":
                if nargs == 2 or n.args[2].t.is_string:
                    derivs = [2]
            elif name == "trace":
                derivs = [1, 2]
            elif name in ("noise", "pnoise") and n.args and n.args[0].t.is_string:
                a0 = n.args[0]
                if a0.kind != "lit" or a0.value == "gabor":
                    k = 2
                    for a in n.args[1:]:
                        if a.t.is_string:
                            break
                        derivs.append(k)
                        k += 1
        self.emit("closure" if rt.is_closure else name, args, "".join(rw), derivs=derivs)
        # write-back for indexed output args
        for w in writes:
            a = n.args[w]
            if a.kind == "index":
                self.store_index(a, argsyms[w])
        return dest

    def cg_usercall(self, n, f, dest):
        fn = f.node
        if fn.body is None:
            raise CompileError("function %s declared but has no body" % f.name)
        actuals = []
        writeback = []
        for a, fm in zip(n.args, fn.formals):
            s = self.cg(a)
            if not fm.out:
                s = self.coerce(s, fm.type)
            elif a.kind == "index":
                writeback.append((a, s))
            actuals.append(s)
        self.scopes.append({})
        saved_aliases = []
        for fm, s in zip(fn.formals, actuals):
            fs = Sym(fm.name, fm.type, "local")
            fs.alias = s
            self.scopes[-1][fm.name] = fs
        op = self.emit("functioncall", [self.const(TSTRING, f.name)], "r")
        self.func_stack.append((f, dest))
        self.scopes.append({})
        self.cg_stmts(fn.body.stmts)
        self.scopes.pop()
        self.func_stack.pop()
        self.scopes.pop()
        self.ops[op].jumps = [self.next_label()]
        for a, s in writeback:
            self.store_index(a, s)
        return dest

    # -- statements ---------------------------------------------------------
    def cg_stmts(self, stmts):
        for s in stmts:
            self.cg_stmt(s)

    def cg_stmt(self, s):
        k = s.kind
        if k == "block":
            if getattr(s, "noscope", False):
                self.cg_stmts(s.stmts)
            else:
                self.scopes.append({})
                self.cg_stmts(s.stmts)
                self.scopes.pop()
        elif k == "expr":
            self.tc(s.e)
            self.cg(s.e)
        elif k == "vardecl":
            self.cg_vardecl(s)
        elif k == "funcdecl":
            self.declare_function(s.f, local=True)
        elif k == "if":
            self.tc(s.cond)
            c = self.cg_int(s.cond)
            ifop = self.emit("if", [c], "r")
            self.scopes.append({})
            self.cg_stmt(s.a)
            self.scopes.pop()
            fl = self.next_label()
            if s.b is not None:
                self.scopes.append({})
                self.cg_stmt(s.b)
                self.scopes.pop()
            self.ops[ifop].jumps = [fl, self.next_label()]
        elif k == "loop":
            lop = self.emit(s.op, [], "")
            self.scopes.append({})
            if s.init is not None:
                self.cg_stmt(s.init)
            cl = self.next_label()
            if s.cond is not None:
                self.tc(s.cond)
                c = self.cg_int(s.cond, True)
            else:
                c = self.const(TINT, 1)
            self.ops[lop].args = [c.deref()]
            self.ops[lop].rw = "r"
            bl = self.next_label()
            self.cg_stmt(s.body)
            il = self.next_label()
            if s.iter is not None:
                self.tc(s.iter)
                self.cg(s.iter)
            self.scopes.pop()
            self.ops[lop].jumps = [cl, bl, il, self.next_label()]
        elif k == "loopmod":
            self.emit(s.op, [], "")
        elif k == "return":
            if not self.func_stack:
                self.emit("exit", [], "")
                return
            f, dest = self.func_stack[-1]
            if s.e is not None:
                self.tc(s.e, f.rtype)
                r = self.cg(s.e, dest)
                if r is not dest:
                    if r.symtype == "const" and r.t.is_int and dest.t.is_float_based:
                        r = self.const(TFLOAT, float(r.value[0]))
                    self.emit("assign", [dest, r])
            self.emit("return", [], "")
        else:
            raise CompileError("unknown statement %s" % k)

    def cg_vardecl(self, s):
        t = s.type
        if t.arr < 0 and s.init is not None and s.init.kind == "initlist":
            t = T(t.base, len(s.init.items))
        sym = self.declare(s.name, t, "local")
        if s.init is None:
            return
        if s.init.kind == "initlist":
            self.tc_initlist_as(s.init, t)
            self.cg_initlist_into(s.init, sym)
            return
        rt = self.tc(s.init, t)
        if not assignable(t, rt):
            raise CompileError("cannot initialize %s %s with %s" % (t, s.name, rt))
        r = self.cg(s.init, sym)
        if r is not sym:
            if r.symtype == "const" and r.t.is_int and t.is_float_based and not t.is_array:
                r = self.const(TFLOAT, float(r.value[0]))
            self.emit("assign", [sym, r])

    def declare_function(self, fn, local=False):
        builtin = any(m[1] == "builtin" for m in fn.meta)
        f = Func(fn.name, fn.rtype, [p.type for p in fn.formals], builtin, fn)
        self.add_func(f)

    # -- shader -------------------------------------------------------------
    def literal_values(self, n, t):
        """Return python list of literal default values for a param, or None
        if the default needs init ops."""
        def lit(x, want):
            if x.kind == "lit":
                if want.base == "string":
                    return [x.value] if x.t.is_string else None
                if x.t.is_string:
                    return None
                if want.base == "int":
                    return [int(x.value)] if x.t.is_int else None
                if want.base == "matrix":
                    v = float(x.value)
                    return [v if (i % 5 == 0) else 0.0 for i in range(16)]
                return [float(x.value)] * want.nfloats()
            if x.kind == "ctor" and x.type.base == want.base and all(
                    a.kind == "lit" and not a.t.is_string for a in x.args):
                vals = [float(a.value) for a in x.args]
                if want.base in TRIPLES:
                    if len(vals) == 1:
                        return vals * 3
                    if len(vals) == 3:
                        return vals
                if want.base == "matrix":
                    if len(vals) == 16:
                        return vals
                    if len(vals) == 1:
                        return [vals[0] if (i % 5 == 0) else 0.0 for i in range(16)]
            return None
        if n is None:
            return None
        if t.is_array:
            if n.kind != "initlist":
                return None
            out = []
            for it in n.items:
                v = lit(it, t.elem())
                if v is None:
                    return None
                out += v
            return out
        return lit(n, t)

    def compile_shader(self, sh):
        self.scopes.append({})
        params = []
        for p in sh.params:
            t = p.type
            unsized = t.arr < 0
            if t.arr < 0 and p.init is not None and p.init.kind == "initlist":
                t = T(t.base, len(p.init.items))
            s = self.declare(p.name, t, "oparam" if p.out else "param")
            s.unsized = unsized      # written as "type[]": the instance value or a connection sets the length
            s.meta = p.meta
            params.append((p, s, t))
        for p, s, t in params:
            vals = self.literal_values(p.init, t)
            if vals is not None:
                s.value = vals
            else:
                n = t.nfloats() * max(1, t.arr)
                s.value = [""] * n if t.base == "string" else (
                    [0] * n if t.base == "int" else [0.0] * n)
                if p.init is not None:
                    s.initexpr = True
                    self.method = s.name
                    if p.init.kind == "initlist":
                        self.tc_initlist_as(p.init, t)
                        self.cg_initlist_into(p.init, s)
                    else:
                        self.tc(p.init, t)
                        r = self.cg(p.init, s)
                        if r is not s:
                            self.emit("assign", [s, r])
        self.method = "___main___"
        self.scopes.append({})
        self.cg_stmts(sh.body.stmts)
        self.scopes.pop()
        self.scopes.pop()
        return self.write_oso(sh)

    # -- writer -------------------------------------------------------------
    def write_oso(self, sh):
        # read/write ranges
        rd, wr = {}, {}
        for i, op in enumerate(self.ops):
            for a, c in zip(op.args, op.rw):
                if c in "rW":
                    lo, hi = rd.get(a, (i, i))
                    rd[a] = (min(lo, i), max(hi, i))
                if c in "wW":
                    lo, hi = wr.get(a, (i, i))
                    wr[a] = (min(lo, i), max(hi, i))
        out = ["OpenShadingLanguage 1.00", "# Compiled by mini_oslc (osl-b200 fixture compiler)",
               "# options: "]
        def metahint(mt, mname, mval):
            if getattr(mval, "kind", None) != "lit":
                return ""
            v = mval.value
            if isinstance(v, str):
                v = '"%s"' % v.replace("\\", "\\\\").replace('"', '\\"')
            elif mt.base == "int":
                v = str(int(v))
            else:
                v = fmt_float(v)
            return " %%meta{%s,%s,%s}" % (mt.base, mname, v)

        hdr = "%s %s" % (sh.stype, sh.name)
        if getattr(self, "emit_metadata", False):
            hdr += "\t" + "".join(metahint(*m) for m in (sh.meta or [])).strip()
        out.append(hdr.rstrip())

        def fmtvals(s):
            if s.value is None:
                return ""
            if s.t.base == "string":
                return " ".join('"%s"' % v.replace("\\", "\\\\").replace('"', '\\"')
                                .replace("\n", "\\n").replace("\t", "\\t") for v in s.value)
            if s.t.base == "int":
                return " ".join(str(int(v)) for v in s.value)
            return " ".join(fmt_float(v) for v in s.value)

        def symline(s):
            tname = repr(s.t)
            if getattr(s, "unsized", False):
                tname = "%s[]" % s.t.base
            line = "%s\t%s\t%s" % (s.symtype, tname, s.name)
            if s.symtype == "const":
                line += "\t" + fmtvals(s) + "\t"
            elif s.symtype in ("param", "oparam"):
                line += "\t" + fmtvals(s) + "\t"
            r = rd.get(s, (2147483647, -1))
            w = wr.get(s, (2147483647, -1))
            line += "\t%%read{%d,%d} %%write{%d,%d}" % (r[0], r[1], w[0], w[1])
            if s.initexpr:
                line += " %initexpr"
            # parameter metadata the runtime acts on: [[ int lockgeom = 0 ]] marks an
            # interpolated (userdata-bound) parameter (oslc writes %meta{type,name,value})
            for mt, mname, mval in (getattr(s, "meta", None) or []):
                if mname == "lockgeom" and getattr(mval, "kind", None) == "lit":
                    line += " %%meta{int,lockgeom,%d}" % int(mval.value)
                elif getattr(self, "emit_metadata", False):
                    line += metahint(mt, mname, mval)
            return line

        for s in self.syms:
            if s.symtype in ("param", "oparam"):
                out.append(symline(s))
        for s in self.syms:
            if s.symtype in ("local", "temp", "global", "const"):
                if s in rd or s in wr:
                    out.append(symline(s))
        last = None
        for op in self.ops:
            if op.method != last:
                out.append("code %s" % op.method)
                last = op.method
            line = "\t" + op.name
            if op.args:
                line += "\t\t" if len(op.name) < 8 else "\t"
            line += "".join(a.name + " " for a in op.args)
            line += "".join("%d " % j for j in op.jumps)
            hints = []
            if op.args:
                hints.append('%%argrw{"%s"}' % op.rw)
            if op.derivs:
                hints.append("%%argderivs{%s}" % ",".join(map(str, op.derivs)))
            if hints:
                line += "\t" + " ".join(hints)
            out.append(line)
        if last != "___main___":
            out.append("code ___main___")
        out.append("\tend")
        return "\n".join(out) + "\n"


def compile_osl(path, include_dirs=(), defines=None, stdosl=None, source=None, metadata=False):
    """Compile an .osl file (or `source` text) and return .oso text.  metadata: write every literal
    [[ ... ]] entry of the shader and its parameters as %meta{type,name,value} hints (oslc does; the
    fixtures only carry the one the runtime acts on, lockgeom)."""
    inc = list(include_dirs)
    pp = Preprocessor(inc, defines)
    text = ""
    if stdosl is None:
        for d in inc:
            if os.path.exists(os.path.join(d, "stdosl.h")):
                stdosl = os.path.join(d, "stdosl.h")
                break
    if stdosl:
        with open(stdosl) as f:
            text += pp.run(f.read(), os.path.dirname(stdosl)) + "\n"
    if source is None:
        with open(path) as f:
            source = f.read()
    curdir = os.path.dirname(os.path.abspath(path)) if path else "."
    text += pp.run(source, curdir)
    funcs, shader = Parser(lex(text)).parse_file()
    if shader is None:
        raise CompileError("no shader found in %s" % path)
    c = Compiler()
    c.emit_metadata = metadata
    for fn in funcs:
        c.declare_function(fn)
    return c.compile_shader(shader)


def main(argv):
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("osl")
    ap.add_argument("-o", dest="out")
    ap.add_argument("-I", dest="inc", action="append", default=[])
    ap.add_argument("-D", dest="defs", action="append", default=[])
    a = ap.parse_args(argv)
    defs = {}
    for d in a.defs:
        k, _, v = d.partition("=")
        defs[k] = (None, v or "1")
    oso = compile_osl(a.osl, a.inc, defs)
    out = a.out or os.path.splitext(os.path.basename(a.osl))[0] + ".oso"
    with open(out, "w") as f:
        f.write(oso)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
