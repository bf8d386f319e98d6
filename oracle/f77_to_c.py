#!/usr/bin/env python3
"""f77_to_c.py -- TEST INFRASTRUCTURE.  Mechanical translator from the fixed-form Fortran 77 subset the SOS-ABS
reference is written in to C, so that the reference's OWN source of the hot-path routines can be compiled and run in
an image that has no Fortran compiler (oracle/build_ref.py drives it; outputs go to oracle/_ref/, never into the
repository history).  It exists to pin oracle/sos_oracle.c: every restated routine is compared bit for bit with the
translation of the routine it restates (tests/test_oracle_vs_reference.py); tests/test_f77_translator.py checks the
translator's own semantics on tiny routines.

What is implemented is what those routines use, with Fortran's own semantics where they differ from C's:
  * cpp-style #include / object-like #define (inc/SOS.h), PARAMETER, fixed-form continuation and tab form, labels,
    '!' comments, statements beyond column 72;
  * INTEGER*4 / REAL / DOUBLE PRECISION / LOGICAL / CHARACTER*n scalars, arrays with arbitrary bounds (column-major),
    IMPLICIT NONE or implicit typing (default and IMPLICIT <type> (ranges)), DIMENSION; all dummy arguments by reference
    with hidden CHARACTER lengths after the argument list (gfortran ABI); local arrays static (gfortran places them in
    zeroed .bss);
  * expression typing by Fortran's rules: REAL*4 literals and all-REAL*4 sub-expressions are evaluated in single
    precision, mixed operands are promoted operand by operand, integer division truncates, x**n with an integer n is
    libgcc's __powidf2 / __powisf2 multiplication chain, x**y otherwise pow / powf; generic and specific intrinsics (DINT
    truncates, DNINT rounds halves away from zero); dotted operators in either case;
  * list-directed READ of numbers and character tokens from text files opened in the routine (a new record per statement, further
    records as the list needs them, ERR= / END=) and from CHARACTER variables / substrings (internal files), READ(u,'(a)') of a
    whole record, `A` / `Aw` output of CHARACTER items, Hollerith and quoted literals in FORMATs, internal WRITE with (Iw) / (Aw)
    into variables and substrings, OPEN with a concatenated file name, GETENV, one-dimensional CHARACTER arrays, character
    constants and PARAMETER constants as actual arguments, blanks inside dotted operators (". AND."); the error messages the
    reference writes on unit 6 are printed to stderr when F77_TO_C_MESSAGES is set (debugging aid);
  * DATA statements for whole one-dimensional arrays, implied DO lists from 1 and scalars (repeat counts n*c), rewritten as
    assignments executed at every call (tables of constants that the routine never modifies);
  * DO loops with the iteration count fixed at entry (labelled, shared terminal labels, ENDDO), DO WHILE, block and
    logical IF, GOTO, CONTINUE, CALL (temporaries for expression arguments), RETURN;
  * I/O as gfortran does it: unformatted sequential records with 4-byte markers (implied DO lists, whole arrays, ERR=,
    END=, IOSTAT=), formatted records on files opened in the routine, also through a computed unit number (nX, Iw, Fw.d,
    Ew.d, Dw.d, repeat groups, format reversion), internal WRITE of an integer with '(Iw)' into a CHARACTER variable, OPEN (STATUS OLD / NEW / UNKNOWN), CLOSE (STATUS='DELETE'), REWIND, INQUIRE(FILE=, EXIST=);
    trace / message WRITEs (list-directed or to units not opened in the routine) are dropped, formatted WRITEs of character
    items (trace tables) abort if they are ever reached;
  * CHARACTER: comparison with blank padding, assignment, substrings, //, INDEX, CALL SYSTEM.
Calls of routines that are not among the translated files abort at run time (they are outside the hot path).
Anything else raises Unsupported and the routine is skipped (reported in oracle/_ref/translation_report.txt).
"""
import re
import sys


class Unsupported(Exception):
    pass


# ----------------------------------------------------------------------------------------------- source lines
def read_defines(path):
    defs = {}
    for ln in open(path, encoding="latin-1"):
        m = re.match(r"#define\s+(\w+)\s+(.*?)\s*$", ln)
        if m:
            defs[m.group(1)] = m.group(2)
    return defs


def strip_comment(s):
    out, q = [], None
    for ch in s:
        if q is None and ch in "'\"":
            q = ch
        elif q is not None and ch == q:
            q = None
        if ch == "!" and q is None:
            break
        out.append(ch)
    return "".join(out).rstrip()


def logical_lines(path, defines):
    """Yields (label, statement_text_upper_outside_strings, first_physical_line_number)."""
    stmts = []
    for no, raw in enumerate(open(path, encoding="latin-1"), 1):
        ln = raw.rstrip("\n").rstrip("\r")
        if not ln.strip():
            continue
        if ln[0] in "Cc*!":
            continue
        if ln.startswith("#"):
            m = re.match(r"#define\s+(\w+)\s+(.*?)\s*$", ln)
            if m:
                defines[m.group(1)] = m.group(2)
            continue
        ti = ln.find("\t")
        if 0 <= ti < 6:                                           # gfortran: a tab in columns 1-6 jumps to the statement field
            lab, rest = ln[:ti], ln[ti + 1:]
            if rest[:1] in tuple("123456789"):
                ln = "     " + rest[0] + rest[1:]
            else:
                ln = lab.ljust(5)[:5] + " " + rest
        ln = ln.replace("\t", " ")
        head = ln[:6].ljust(6)
        body = strip_comment(ln[6:72] if False else ln[6:])      # the reference exceeds column 72 freely (gfortran -ffixed-line-length-none)
        if head[:5].strip().startswith("!"):
            continue
        if head[5] not in " 0" and head[:5].strip() == "":
            if not stmts:
                raise Unsupported("continuation without a statement at line %d" % no)
            stmts[-1][1] += " " + body.strip()
            continue
        if not body.strip():
            continue
        stmts.append([head[:5].strip(), body.strip(), no])
    for lab, txt, no in stmts:
        yield lab, txt, no


TOKEN_RE = re.compile(r"""\s*(?:
    (?P<str>'(?:[^']|'')*'|"[^"]*") |
    (?P<dotop>\.\s*(?i:EQ|NE|LT|LE|GT|GE|AND|OR|NOT|TRUE|FALSE|EQV|NEQV)\s*\.) |
    (?P<num>(?:\d+\.?\d*|\.\d+)(?:[EDed][+-]?\d+)?) |
    (?P<id>[A-Za-z_][A-Za-z_0-9]*) |
    (?P<op>\*\*|//|[-+*/(),=:])
)""", re.X)


def tokenize(text, defines, depth=0):
    toks, pos = [], 0
    text = text.rstrip()
    while pos < len(text):
        m = TOKEN_RE.match(text, pos)
        if not m:
            raise Unsupported("cannot tokenize %r" % text[pos:pos + 20])
        pos = m.end()
        if m.group("str"):
            v = m.group("str")
            if v.startswith('"'):
                v = "'" + v[1:-1].replace("'", "''") + "'"
            toks.append(("str", v))
        elif m.group("dotop"):
            toks.append(("op", re.sub(r"\s+", "", m.group("dotop")).upper()))       # blanks are not significant in fixed form: ". AND."
        elif m.group("num"):
            s = m.group("num").upper()
            # "1.EQ." style: a trailing '.' followed by a dotted operator belongs to the operator
            if s.endswith(".") and re.match(r"(EQ|NE|LT|LE|GT|GE|AND|OR)\.", text[pos:pos + 4].upper()):
                s = s[:-1]
                pos -= 1
            toks.append(("num", s))
        elif m.group("id"):
            name = m.group("id")
            up = name.upper()
            key = name if name in defines else (up if up in defines else None)
            if key is not None and depth < 8:
                toks.extend(tokenize(defines[key], defines, depth + 1))
            else:
                toks.append(("id", up))
        else:
            toks.append(("op", m.group("op")))
    return toks


# ----------------------------------------------------------------------------------------------- expressions
RANK = {"i": 0, "r": 1, "d": 2}
CT = {"i": "int", "r": "float", "d": "double", "l": "int"}

INTRINSICS = {
    # name: (kind, ...)   kind 'd1': double function of one arg; 'g1': generic by argument type; 'conv': conversion
    "DEXP": ("d1", "exp"), "DSQRT": ("d1", "sqrt"), "DCOS": ("d1", "cos"), "DSIN": ("d1", "sin"), "DTAN": ("d1", "tan"),
    "DACOS": ("d1", "acos"), "DASIN": ("d1", "asin"), "DATAN": ("d1", "atan"), "DLOG": ("d1", "log"), "DLOG10": ("d1", "log10"),
    "DABS": ("d1", "fabs"), "DINT": ("d1", "trunc"), "DNINT": ("d1", "round"),      # DNINT: halves away from zero, as C round()
    "EXP": ("g1", "exp"), "SQRT": ("g1", "sqrt"), "COS": ("g1", "cos"), "SIN": ("g1", "sin"), "TAN": ("g1", "tan"),
    "ACOS": ("g1", "acos"), "ASIN": ("g1", "asin"), "ATAN": ("g1", "atan"), "LOG": ("g1", "log"), "ALOG": ("g1", "log"),
    "ABS": ("abs",), "IABS": ("abs",),
    "DBLE": ("conv", "d"), "DFLOAT": ("conv", "d"), "FLOAT": ("conv", "r"), "REAL": ("conv", "r"), "SNGL": ("conv", "r"),
    "INT": ("conv", "i"), "IDINT": ("conv", "i"), "IFIX": ("conv", "i"),
    "MAX": ("minmax", ">"), "MIN": ("minmax", "<"), "DMAX1": ("minmax", ">"), "DMIN1": ("minmax", "<"),
    "AMAX1": ("minmax", ">"), "AMIN1": ("minmax", "<"), "MAX0": ("minmax", ">"), "MIN0": ("minmax", "<"),
    "MOD": ("mod",),
}


def conv(code, frm, to):
    if frm == to:
        return code
    if to == "l" or frm == "l":
        return code
    return "((%s)(%s))" % (CT[to], code)


class Parser:
    def __init__(self, toks, unit):
        self.t, self.p, self.u = toks, 0, unit

    def peek(self):
        return self.t[self.p] if self.p < len(self.t) else ("end", "")

    def take(self, val=None):
        k, v = self.peek()
        if val is not None and v != val:
            raise Unsupported("expected %r, found %r" % (val, v))
        self.p += 1
        return k, v

    def at(self, val):
        return self.peek()[1] == val and self.peek()[0] == "op"

    # precedence climbing
    def expr(self):
        return self.p_or()

    def p_or(self):
        a = self.p_or1()
        while self.at(".EQV.") or self.at(".NEQV."):
            op = self.take()[1]
            b = self.p_or1()
            a = ("((!!(%s)) %s (!!(%s)))" % (a[0], "==" if op == ".EQV." else "!=", b[0]), "l")
        return a

    def p_or1(self):
        a = self.p_and()
        while self.at(".OR."):
            self.take()
            b = self.p_and()
            a = ("(%s || %s)" % (a[0], b[0]), "l")
        return a

    def p_and(self):
        a = self.p_not()
        while self.at(".AND."):
            self.take()
            b = self.p_not()
            a = ("(%s && %s)" % (a[0], b[0]), "l")
        return a

    def p_not(self):
        if self.at(".NOT."):
            self.take()
            a = self.p_not()
            return ("(!%s)" % a[0], "l")
        return self.p_rel()

    def p_rel(self):
        a = self.p_add()
        ops = {".EQ.": "==", ".NE.": "!=", ".LT.": "<", ".LE.": "<=", ".GT.": ">", ".GE.": ">="}
        if self.peek()[0] == "op" and self.peek()[1] in ops:
            op = ops[self.take()[1]]
            b = self.p_add()
            if a[1] == "c" or b[1] == "c":
                if a[1] != "c" or b[1] != "c" or op not in ("==", "!="):
                    raise Unsupported("character comparison")
                return ("(f77_cmp(%s, %s) %s 0)" % (a[0], b[0], op), "l")
            t = a[1] if RANK[a[1]] >= RANK[b[1]] else b[1]
            return ("(%s %s %s)" % (conv(a[0], a[1], t), op, conv(b[0], b[1], t)), "l")
        return a

    def p_add(self):
        if self.at("-") or self.at("+"):
            op = self.take()[1]
            a = self.p_mul()
            a = ("(%s%s)" % (op, a[0]), a[1])
        else:
            a = self.p_mul()
        while self.at("+") or self.at("-"):
            op = self.take()[1]
            b = self.p_mul()
            a = self.arith(a, op, b)
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.at("*") or self.at("/"):
            op = self.take()[1]
            b = self.p_pow()
            a = self.arith(a, op, b)
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.at("**"):
            self.take()
            if self.at("-") or self.at("+"):
                raise Unsupported("signed exponent without parentheses")
            b = self.p_pow()                                     # right associative
            if b[1] == "i":
                f = {"i": "f77_powi_i", "r": "f77_powi_f", "d": "f77_powi_d"}[a[1]]
                return ("%s(%s, %s)" % (f, a[0], b[0]), a[1])
            t = a[1] if RANK[a[1]] >= RANK[b[1]] else b[1]
            if t == "i":
                t = b[1]
            f = "powf" if t == "r" else "pow"
            return ("%s(%s, %s)" % (f, conv(a[0], a[1], t), conv(b[0], b[1], t)), t)
        return a

    def arith(self, a, op, b):
        if a[1] not in RANK or b[1] not in RANK:
            raise Unsupported("arithmetic on non-numeric operands")
        t = a[1] if RANK[a[1]] >= RANK[b[1]] else b[1]
        return ("(%s %s %s)" % (conv(a[0], a[1], t), op, conv(b[0], b[1], t)), t)

    def p_primary(self):
        k, v = self.take()
        if k == "num":
            if re.search(r"D", v):
                return ("%s" % v.replace("D", "e"), "d")
            if "." in v or "E" in v:
                s = v
                if s.endswith("."):
                    s += "0"
                if s.startswith("."):
                    s = "0" + s
                return ("%sf" % s, "r")
            return (v, "i")
        if k == "op" and v == "(":
            a = self.expr()
            self.take(")")
            return ("(%s)" % a[0], a[1])
        if k == "op" and v in (".TRUE.", ".FALSE."):
            return ("1" if v == ".TRUE." else "0", "l")
        if k == "str":
            lit = v[1:-1].replace("''", "'")
            return ('"%s", %d' % (lit.replace("\\", "\\\\").replace('"', '\\"'), len(lit)), "c")
        if k == "id":
            if self.at("(") and v in self.u.vars and self.u.vars[v]["type"] == "c" and len(self.u.vars[v]["dims"]) == 1:
                self.take("(")                                      # element of a one-dimensional CHARACTER array
                ix = self.expr()
                self.take(")")
                lo = self.u.cstr(self.u.vars[v]["dims"][0][0], "i")
                ln = self.u.cstr(self.u.vars[v]["clen"], "i")
                return ("%s + (size_t)((%s) - (%s)) * (size_t)(%s), (size_t)(%s)" % (self.u.cname(v), ix[0], lo, ln, ln), "c")
            if self.at("(") and v in self.u.vars and self.u.vars[v]["type"] == "c" and not self.u.vars[v]["dims"]:
                self.take("(")
                lo = self.expr()
                self.take(":")
                hi = self.expr()
                self.take(")")
                base = self.u.cname(v)
                return ("%s + ((%s) - 1), (size_t)((%s) - (%s) + 1)" % (base, lo[0], hi[0], lo[0]), "c")
            if self.at("("):
                self.take("(")
                args = []
                if not self.at(")"):
                    args.append(self.expr())
                    while self.at(","):
                        self.take()
                        args.append(self.expr())
                self.take(")")
                if v in self.u.vars and self.u.vars[v]["dims"]:
                    return (self.u.array_ref(v, args), self.u.vars[v]["type"])
                if v == "INDEX" and len(args) == 2 and args[0][1] == "c" and args[1][1] == "c":
                    return ("f77_index(%s, %s)" % (args[0][0], args[1][0]), "i")
                return self.intrinsic(v, args)
            if v in self.u.vars:
                if self.u.vars[v]["type"] == "c":
                    if v not in self.u.args:
                        return ("%s, (size_t)(%s)" % (self.u.cname(v), self.u.cstr(self.u.vars[v]["clen"], "i")), "c")
                    return ("%s, len_%s" % (self.u.cname(v), self.u.cname(v)), "c")
                return (self.u.scalar_ref(v), self.u.vars[v]["type"])
            if not self.u.implicit_none and not self.at("("):
                # Fortran implicit typing: I-N integer, otherwise REAL
                self.u.vars[v] = {"type": self.u.implicit[v[0]], "dims": [], "clen": "1"}
                return (self.u.scalar_ref(v), self.u.vars[v]["type"])
            raise Unsupported("undeclared name %s" % v)
        raise Unsupported("unexpected token %r" % (v,))

    def intrinsic(self, name, args):
        if name not in INTRINSICS:
            raise Unsupported("function %s" % name)
        spec = INTRINSICS[name]
        kind = spec[0]
        if kind == "d1":
            return ("%s(%s)" % (spec[1], conv(args[0][0], args[0][1], "d")), "d")
        if kind == "g1":
            t = args[0][1]
            if t == "i":
                raise Unsupported("%s of an integer" % name)
            return ("%s%s(%s)" % (spec[1], "f" if t == "r" else "", args[0][0]), t)
        if kind == "abs":
            t = args[0][1]
            return ("%s(%s)" % ({"i": "abs", "r": "fabsf", "d": "fabs"}[t], args[0][0]), t)
        if kind == "conv":
            to = spec[1]
            return (conv(args[0][0], args[0][1], to), to)
        if kind == "minmax":
            t = max((a[1] for a in args), key=lambda x: RANK[x])
            code = conv(args[0][0], args[0][1], t)
            for a in args[1:]:
                b = conv(a[0], a[1], t)
                code = "((%s) %s (%s) ? (%s) : (%s))" % (code, spec[1], b, code, b)
            return (code, t)
        if kind == "mod":
            a, b = args
            if a[1] == "i" and b[1] == "i":
                return ("((%s) %% (%s))" % (a[0], b[0]), "i")
            t = a[1] if RANK[a[1]] >= RANK[b[1]] else b[1]
            return ("%s(%s, %s)" % ("fmodf" if t == "r" else "fmod", conv(a[0], a[1], t), conv(b[0], b[1], t)), t)
        raise Unsupported(name)


# ----------------------------------------------------------------------------------------------- program units
TYPE_RE = re.compile(r"^(DOUBLE\s*PRECISION|REAL\s*\*\s*8|REAL\s*\*\s*4|REAL|INTEGER\s*\*\s*4|INTEGER\s*\*\s*2|INTEGER|LOGICAL\s*\*\s*4|LOGICAL|"
                     r"CHARACTER\s*\*\s*\(?\s*[\w]+\s*\)?|CHARACTER)\s*(.*)$", re.I)


def split_top(s, sep=","):
    out, depth, cur, q = [], 0, [], False
    for ch in s:
        if ch == "'":
            q = not q
        if not q:
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == sep and depth == 0:
                out.append("".join(cur).strip())
                cur = []
                continue
        cur.append(ch)
    if "".join(cur).strip():
        out.append("".join(cur).strip())
    return out


class Unit:
    def __init__(self, name, args, defines):
        self.name, self.args, self.defines = name, args, defines
        self.vars = {}
        self.body = []                 # C lines
        self.labels_used = set()
        self.stubs = set()
        self.params = set()                         # names defined by PARAMETER statements
        self.formats = {}
        self.implicit_none = False
        self.implicit = {ch: ("i" if ch in "IJKLMN" else "r") for ch in "ABCDEFGHIJKLMNOPQRSTUVWXYZ"}
        self.text_units = set()
        self.tmp = 0

    def cname(self, v):
        return "v_" + v.lower()

    def scalar_ref(self, v):
        return ("(*%s)" % self.cname(v)) if v in self.args else self.cname(v)

    def dim_exprs(self, v):
        out = []
        for lo, hi in self.vars[v]["dims"]:
            out.append((self.cstr(lo, "i"), None if hi == "*" else self.cstr(hi, "i")))
        return out

    def array_ref(self, v, args):
        dims = self.dim_exprs(v)
        if len(args) != len(dims):
            raise Unsupported("rank mismatch on %s" % v)
        idx, stride = None, None
        for (a, t), (lo, hi) in zip(args, dims):
            if t != "i":
                raise Unsupported("non-integer subscript")
            term = "((%s) - (%s))" % (a, lo)
            if idx is None:
                idx = term
            else:
                idx = "%s + (%s) * %s" % (idx, stride, term)
            if hi is not None:
                ext = "((%s) - (%s) + 1)" % (hi, lo)
                stride = ext if stride is None else "(%s) * %s" % (stride, ext)
        return "%s[%s]" % (self.cname(v), idx)

    def cexpr(self, text_or_toks):
        """(C code, Fortran type) of an expression"""
        toks = tokenize(text_or_toks, self.defines) if isinstance(text_or_toks, str) else text_or_toks
        p = Parser(toks, self)
        e = p.expr()
        if p.p != len(toks):
            raise Unsupported("trailing tokens in expression: %r" % (toks[p.p:],))
        return e

    def cstr(self, text_or_toks, want):
        """C code of an expression converted to the wanted type"""
        e = self.cexpr(text_or_toks)
        return conv(e[0], e[1], want)

    def declare(self, ftype, rest):
        ft = re.sub(r"\s+", "", ftype.upper())
        if ft.startswith("DOUBLE") or ft == "REAL*8":
            t = "d"
        elif ft.startswith("REAL"):
            t = "r"
        elif ft.startswith("INTEGER"):
            t = "i"
        elif ft.startswith("LOGICAL"):
            t = "l"
        else:
            t = "c"
        m0 = re.match(r"^CHARACTER\*\(?(\w+)\)?$", ft)
        clen = m0.group(1) if m0 else "1"
        for item in split_top(rest):
            m = re.match(r"^(\w+)\s*(?:\((.*)\))?\s*(?:\*\s*\(?\s*\w+\s*\)?)?$", item.strip())
            if not m:
                raise Unsupported("declaration %r" % item)
            name = m.group(1).upper()
            dims = []
            if m.group(2) is not None:
                for d in split_top(m.group(2)):
                    if ":" in d:
                        lo, hi = [x.strip() for x in split_top(d, ":")]
                    else:
                        lo, hi = "1", d.strip()
                    dims.append((lo, hi))
            self.vars[name] = {"type": t, "dims": dims, "clen": clen}


def expand_data(exe):
    """DATA statements of the two forms the reference uses -- `DATA A/ c1, c2, ... /` (a scalar, or a whole one-dimensional array
    with lower bound 1) and `DATA (A(I), I = 1, N)/ c1, ... /` -- rewritten as assignments A(k) = c_k at the place of the statement
    (repeat counts n*c expanded).  The assignments are executed at every call instead of once at load time, which is the same thing
    for variables the routine never modifies afterwards (the case of every DATA statement of the reference: tables of constants);
    the constants keep their own type (a REAL*4 constant stored in a DOUBLE PRECISION element is converted, as DATA does)."""
    out = []
    for lab, txt, no in exe:
        m = re.match(r"^DATA\s+(.*)$", txt, re.I | re.S)
        if not m or re.match(r"^DATA\w*\s*(\(.*\))?\s*=", txt, re.I):
            out.append((lab, txt, no))
            continue
        body = m.group(1).strip()
        m1 = re.match(r"^\(\s*(\w+)\s*\(\s*(\w+)\s*\)\s*,\s*(\w+)\s*=\s*1\s*,\s*[^)]*\)\s*/(.*)/\s*$", body, re.S)
        m2 = re.match(r"^(\w+)\s*/(.*)/\s*$", body, re.S)
        if m1 and m1.group(2).upper() == m1.group(3).upper():
            var, vals = m1.group(1), m1.group(4)
        elif m2:
            var, vals = m2.group(1), m2.group(2)
        else:
            raise Unsupported("DATA form: %s" % txt[:60])
        consts = []
        for c in split_top(vals):
            c = c.strip()
            r = re.match(r"^(\d+)\s*\*\s*(.*)$", c)
            consts += [r.group(2)] * int(r.group(1)) if r else [c]
        if lab:
            out.append((lab, "CONTINUE", no))
        if len(consts) == 1 and m2 and not m1:
            out.append(("", "%s = %s" % (var, consts[0]), no))               # a scalar (a one-element array is not used by the reference)
        else:
            for k, c in enumerate(consts):
                out.append(("", "%s(%d) = %s" % (var, k + 1, c), no))
    return out


def translate_unit(name, args, stmts, defines, known_subs):
    u = Unit(name, args, defines)
    out = []
    # pass 1: declarations
    exe = []
    pending_dims = []
    for lab, txt, no in stmts:
        up = txt.upper()
        m = TYPE_RE.match(txt)
        if m and not re.match(r"^\w+\s*(\(.*\))?\s*=", txt):
            u.declare(m.group(1), m.group(2))
            continue
        if up.startswith("IMPLICIT"):
            if re.match(r"^IMPLICIT\s+NONE", up):
                u.implicit_none = True
                continue
            m2 = re.match(r"^IMPLICIT\s+(DOUBLE\s*PRECISION|REAL\s*\*\s*8|REAL|INTEGER(?:\s*\*\s*4)?|LOGICAL)\s*\((.*)\)\s*$", up)
            if not m2:
                raise Unsupported("IMPLICIT form: %s" % up)
            tt = "d" if (m2.group(1).startswith("DOUBLE") or m2.group(1).replace(" ", "") == "REAL*8") else ("r" if m2.group(1).startswith("REAL") else ("i" if m2.group(1).startswith("INTEGER") else "l"))
            for rng in m2.group(2).split(","):
                a, _, b = rng.strip().partition("-")
                for o_ in range(ord(a.strip()), ord((b or a).strip()) + 1):
                    u.implicit[chr(o_)] = tt
            continue
        m2 = re.match(r"^PARAMETER\s*\((.*)\)\s*$", txt, re.I | re.S)
        if m2:
            for item in split_top(m2.group(1)):
                k, v = item.split("=", 1)
                u.defines = dict(u.defines)
                u.defines[k.strip().upper()] = "(" + v.strip() + ")"
                u.params.add(k.strip().upper())
            continue
        m2 = re.match(r"^DIMENSION\s+(.*)$", txt, re.I | re.S)
        if m2:
            pending_dims.append(m2.group(1))
            continue
        exe.append((lab, txt, no))
    exe = expand_data(exe)
    for lab, txt, no in exe:
        m = re.match(r"^FORMAT\s*(\(.*\))\s*$", txt, re.I | re.S)
        if m and lab:
            u.formats[lab] = m.group(1)
        m = re.match(r"^OPEN\s*\((.*)\)\s*$", txt, re.I | re.S)
        if m and "UNFORMATTED" not in m.group(1).upper():
            for c in split_top(m.group(1)):
                if c.upper().replace(" ", "").startswith("UNIT="):
                    u.text_units.add(c.split("=", 1)[1].strip())
                elif "=" not in c:
                    u.text_units.add(c.strip())
    for text in pending_dims:                                     # DIMENSION with implicit typing
        for item in split_top(text):
            nm = item.split("(")[0].strip().upper()
            if nm in u.vars:
                t0 = {"d": "DOUBLE PRECISION", "r": "REAL", "i": "INTEGER", "l": "LOGICAL"}[u.vars[nm]["type"]]
            else:
                t0 = {"d": "DOUBLE PRECISION", "r": "REAL", "i": "INTEGER", "l": "LOGICAL"}[u.implicit[nm[0]]]
            u.declare(t0, item)
    for a in args:
        if a not in u.vars:
            if u.implicit_none:
                raise Unsupported("argument %s has no declaration" % a)
            u.vars[a] = {"type": u.implicit[a[0]], "dims": [], "clen": "1"}
    has_char = any(u.vars[a]["type"] == "c" for a in args)
    # DO-loop bookkeeping
    do_stack = []          # (terminal label or None)
    ind = ["  "]

    def emit(s):
        out.append("".join(ind) + s)

    def new_tmp():
        u.tmp += 1
        return "t%d_" % u.tmp

    def simple_statement(txt):
        """assignment / GOTO / CALL / RETURN / CONTINUE / WRITE as a single C statement string"""
        up = txt.upper().strip()
        if up == "CONTINUE":
            return ";"
        if up == "RETURN":
            return "return;"
        m = re.match(r"^GO\s*TO\s*(\d+)$", up)
        if m:
            u.labels_used.add(m.group(1))
            return "goto L%s;" % m.group(1)
        m = re.match(r"^(WRITE|READ|PRINT)\s*\((.*)$", txt, re.I | re.S)
        if m or up.startswith("PRINT"):
            if up.startswith("PRINT"):
                return "; /* PRINT dropped */"
            kind = m.group(1).upper()
            rest = m.group(2)
            depth, j = 1, 0
            for j, ch in enumerate(rest):
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                    if depth == 0:
                        break
            ctl, items = split_top(rest[:j]), rest[j + 1:].strip()
            pos = [c for c in ctl if "=" not in c]
            kv = dict((c.split("=", 1)[0].strip().upper(), c.split("=", 1)[1].strip()) for c in ctl if "=" in c)
            formatted = len(pos) > 1 or "FMT" in kv
            unit_txt = kv.get("UNIT", pos[0] if pos else "0").strip()
            # internal file: WRITE(CHARVAR,'(Iw)') integer-expression  (SOS_NOM_FICMIE builds the MIE file name this way)
            mi = re.match(r"^'\(\s*([IA])(\d+)\s*\)'$", (pos[1] if len(pos) > 1 else "").strip(), re.I)
            unit_name = re.match(r"^\s*(\w+)", unit_txt)
            if kind == "WRITE" and mi and unit_name and unit_name.group(1).upper() in u.vars and u.vars[unit_name.group(1).upper()]["type"] == "c":
                dst = u.cexpr(unit_txt)                              # a CHARACTER variable or a substring of one
                val = u.cexpr(rest[j + 1:].strip())
                if mi.group(1).upper() == "A":
                    if val[1] != "c":
                        raise Unsupported("internal WRITE: A field with a non-character item")
                    return "{ static char a_[4096]; size_t n_ = f77_cat(a_, 0, %s); if (n_ > %d) n_ = %d; f77_assign(%s, a_, n_); }" % (
                        val[0], int(mi.group(2)), int(mi.group(2)), dst[0])
                if val[1] != "i":
                    raise Unsupported("internal WRITE of a non-integer")
                return "{ char b_[64]; snprintf(b_, sizeof b_, \"%%*d\", %d, (int)(%s)); f77_assign(%s, b_, strlen(b_)); }" % (int(mi.group(2)), val[0], dst[0])
            # READ(u,'(a)') CHARVAR: one record of a text file into a CHARACTER variable
            if kind == "READ" and len(pos) > 1 and re.match(r"^'\(\s*A\s*\)'$", pos[1].strip(), re.I) and unit_txt in u.text_units:
                dst = u.cexpr(rest[j + 1:].strip())
                if dst[1] != "c":
                    raise Unsupported("READ '(a)' into a non-character item")
                err = kv.get("ERR")
                if err:
                    u.labels_used.add(err)
                return "if (!f77_readline(%s, %s)) { %s }" % (u.cstr(unit_txt, "i"), dst[0], "goto L%s;" % err if err else "abort();")
            fmt_lab = kv.get("FMT", pos[1] if len(pos) > 1 else None)
            text_file = formatted and unit_txt in u.text_units and fmt_lab is not None and fmt_lab.strip() in u.formats
            if formatted and not text_file and not re.match(r"^\d+$", unit_txt) and not re.match(r"^\w+$", unit_txt) and u.text_units \
                    and fmt_lab is not None and fmt_lab.strip() in u.formats:
                text_file = True                                   # computed unit, e.g. READ((I+10),555): one of the units opened here
            # list-directed READ from a text file opened here, READ(u,*[,ERR=][,END=]) items, or from a CHARACTER variable / substring
            list_mode = kind == "READ" and fmt_lab is not None and fmt_lab.strip() == "*" and unit_txt in u.text_units
            internal = None
            if kind == "READ" and fmt_lab is not None and fmt_lab.strip() == "*" and not list_mode and unit_name \
                    and unit_name.group(1).upper() in u.vars and u.vars[unit_name.group(1).upper()]["type"] == "c":
                internal = u.cexpr(unit_txt)
                list_mode = True
            if formatted and not text_file and not list_mode:
                if kind == "READ":
                    raise Unsupported("formatted READ on a unit that is not opened here")
                lit = re.match(r"^\s*('(?:[^']|'')*'|\"[^\"]*\")", items)
                if unit_txt == "6" and lit:                        # error messages of the reference: shown on request (debugging aid)
                    txt_ = lit.group(1)[1:-1].replace("\\", "\\\\").replace('"', '\\"')
                    return "if (getenv(\"F77_TO_C_MESSAGES\")) fprintf(stderr, \"%%s\\n\", \"%s\");" % txt_
                return "; /* trace / message WRITE dropped */"
            unit = "100" if internal else u.cstr(unit_txt, "i")
            err = kv.get("ERR")
            fail = "goto L%s;" % err if err else "abort();"
            if err:
                u.labels_used.add(err)
            endl = kv.get("END")
            if endl:
                u.labels_used.add(endl)

            def io_items(text):
                code = []
                for it in split_top(text):
                    it = it.strip()
                    if it.startswith("(") and it.endswith(")") and any(re.match(r"^\w+\s*=", q) for q in split_top(it[1:-1])):
                        parts = split_top(it[1:-1])
                        k = next(i for i, q in enumerate(parts) if re.match(r"^\w+\s*=", q))
                        var, e1 = [x.strip() for x in parts[k].split("=", 1)]
                        e2 = parts[k + 1]
                        e3 = parts[k + 2] if len(parts) > k + 2 else "1"
                        v = u.scalar_ref(var.upper())
                        code.append("for (%s = %s; (%s) > 0 ? %s <= (%s) : %s >= (%s); %s += %s) {" %
                                    (v, u.cstr(e1, "i"), u.cstr(e3, "i"), v, u.cstr(e2, "i"), v, u.cstr(e2, "i"), v, u.cstr(e3, "i")))
                        code += io_items(", ".join(parts[:k]))
                        code.append("}")
                    elif re.match(r"^\w+$", it) and it.upper() in u.vars and u.vars[it.upper()]["dims"]:
                        v = u.vars[it.upper()]                         # whole array, storage order
                        n = " * ".join("((%s) - (%s) + 1)" % (u.cstr(hi, "i"), u.cstr(lo, "i")) for lo, hi in v["dims"])
                        if text_file or list_mode:
                            raise Unsupported("whole-array formatted I/O")
                        if kind == "READ":
                            code.append("if (!f77_ritem(%s, %s, (size_t)(%s) * sizeof(%s))) { %s }" % (unit, u.cname(it.upper()), n, CT[v["type"]], fail))
                        else:
                            code.append("f77_witem(%s, %s, (size_t)(%s) * sizeof(%s));" % (unit, u.cname(it.upper()), n, CT[v["type"]]))
                    else:
                        e = u.cexpr(it)
                        if e[1] == "c" and list_mode:
                            code.append("if (!f77_lread_c(%s, %s)) { %s }" % (unit, e[0], fail))
                            continue
                        if e[1] == "c" and text_file and kind == "WRITE":
                            code.append("f77_fwrite_a(%s, %s);" % (unit, e[0]))
                            continue
                        if e[1] == "c":
                            raise Unsupported("character I/O item")
                        if list_mode:
                            code.append("{ double t_; if (!f77_lread(%s, &t_)) { %s } %s = (%s)t_; }" % (unit, fail, e[0], CT[e[1]]))
                        elif text_file:
                            if kind == "READ":
                                code.append("{ double t_; if (!f77_fread(%s, &t_)) { %s } %s = (%s)t_; }" % (unit, fail, e[0], CT[e[1]]))
                            else:
                                code.append("f77_fwrite(%s, (double)(%s), %d);" % (unit, e[0], 1 if e[1] == "i" else 0))
                        elif kind == "READ":
                            code.append("if (!f77_ritem(%s, &%s, sizeof(%s))) { %s }" % (unit, e[0], CT[e[1]], fail))
                        else:
                            tmp = new_tmp()
                            code.append("{ %s %s = %s; f77_witem(%s, &%s, sizeof(%s)); }" % (CT[e[1]], tmp, e[0], unit, tmp, CT[e[1]]))
                return code
            try:
                body = io_items(items)
            except Unsupported as ex:
                if kind == "WRITE" and text_file and "character I/O item" in str(ex):
                    return "abort(); /* formatted WRITE of character items (trace output): not translated, must not be reached */"
                raise
            if list_mode:                                          # a new record per statement; more records as the list needs them
                at_end = ("goto L%s;" % endl) if endl else fail
                if internal:
                    return "{ f77_ibegin(%s); %s }" % (internal[0], " ".join(body))
                return "{ int s_ = f77_lbegin(%s); if (s_ < 0) { %s } if (!s_) { %s } %s }" % (unit, at_end, fail, " ".join(body))
            if text_file:
                fmt = u.formats[fmt_lab.strip()].replace("\\", "\\\\").replace('"', '\\"')
                if kind == "READ":
                    return "{ if (!f77_fbegin(%s, \"%s\", 1)) { %s } %s }" % (unit, fmt, fail, " ".join(body))
                return "{ if (!f77_fbegin(%s, \"%s\", 0)) { %s } %s if (!f77_fend(%s)) { %s } }" % (unit, fmt, fail, " ".join(body), unit, fail)
            if kind == "READ":
                at_end = ("if (f77_eof(%s)) goto L%s; " % (unit, endl)) if endl else ""
                if "IOSTAT" in kv:
                    ios = u.cexpr(kv["IOSTAT"])[0]
                    return "{ %s = 0; if (f77_eof(%s)) { %s = -1; } else { if (!f77_rbegin(%s)) { %s } %s } }" % (ios, unit, ios, unit, fail, " ".join(body))
                return "{ %sif (!f77_rbegin(%s)) { %s } %s }" % (at_end, unit, fail, " ".join(body))
            return "{ f77_wbegin(%s); %s if (!f77_wend(%s)) { %s } }" % (unit, " ".join(body), unit, fail)
        m = re.match(r"^OPEN\s*\((.*)\)\s*$", txt, re.I | re.S)
        if m:
            kv = {}
            for c in split_top(m.group(1)):
                if "=" in c:
                    kv[c.split("=", 1)[0].strip().upper()] = c.split("=", 1)[1].strip()
                else:
                    kv["UNIT"] = c.strip()
            text = "UNFORMATTED" not in kv.get("FORM", "").upper()
            err = kv.get("ERR")
            if err:
                u.labels_used.add(err)
            st = kv.get("STATUS", "'UNKNOWN'").upper().strip("'")
            flags = {"OLD": 1, "NEW": 2}.get(st, 0) + (4 if text else 0)
            fail = "goto L%s;" % err if err else "abort();"
            ftoks = tokenize(kv["FILE"], u.defines)
            if any(k == "op" and v == "//" for k, v in ftoks):        # FILE = a // b // ...: concatenated into a buffer first
                parts, cur, depth = [], [], 0
                for k, v in ftoks:
                    if k == "op" and v == "(":
                        depth += 1
                    if k == "op" and v == ")":
                        depth -= 1
                    if k == "op" and v == "//" and depth == 0:
                        parts.append(cur)
                        cur = []
                    else:
                        cur.append((k, v))
                parts.append(cur)
                code = "{ static char cat_[4096]; size_t n_ = 0; "
                for pt in parts:
                    e = u.cexpr(pt)
                    if e[1] != "c":
                        raise Unsupported("concatenation of a non-character operand")
                    code += "n_ = f77_cat(cat_, n_, %s); " % e[0]
                return code + "if (!f77_open(%s, cat_, n_, %d)) { %s } }" % (u.cstr(kv["UNIT"], "i"), flags, fail)
            f = u.cexpr(kv["FILE"])
            if f[1] != "c":
                raise Unsupported("OPEN FILE= expression")
            return "if (!f77_open(%s, %s, %d)) { %s }" % (u.cstr(kv["UNIT"], "i"), f[0], flags, fail)
        m = re.match(r"^CLOSE\s*\((.*)\)\s*$", txt, re.I | re.S)
        if m:
            parts = split_top(m.group(1))
            unit = [c for c in parts if "=" not in c or c.upper().replace(" ", "").startswith("UNIT=")][0].split("=")[-1]
            delete = any("DELETE" in c.upper() for c in parts)
            return "f77_close(%s, %d);" % (u.cstr(unit, "i"), 1 if delete else 0)
        m = re.match(r"^REWIND\s*\(?\s*(\w+)\s*\)?\s*$", txt, re.I)
        if m:
            return "f77_rewind(%s);" % u.cstr(m.group(1), "i")
        if re.match(r"^BACKSPACE\b", up):
            raise Unsupported("file I/O: %s" % up[:20])
        m = re.match(r"^INQUIRE\s*\((.*)\)\s*$", txt, re.I | re.S)
        if m:
            kv = dict((c.split("=", 1)[0].strip().upper(), c.split("=", 1)[1].strip()) for c in split_top(m.group(1)) if "=" in c)
            if "FILE" not in kv or "EXIST" not in kv:
                raise Unsupported("INQUIRE form")
            f = u.cexpr(kv["FILE"])
            return "%s = f77_exists(%s);" % (u.cexpr(kv["EXIST"])[0], f[0])
        m = re.match(r"^CALL\s+SYSTEM\s*\((.*)\)\s*$", txt, re.I | re.S)
        if m:
            return "f77_system(%s);" % u.cexpr(m.group(1))[0]
        m = re.match(r"^CALL\s+GETENV\s*\((.*)\)\s*$", txt, re.I | re.S)
        if m:
            a = split_top(m.group(1))
            name, var = u.cexpr(a[0]), u.cexpr(a[1])
            if name[1] != "c" or var[1] != "c" or not name[0].startswith('"'):
                raise Unsupported("GETENV form")
            return "{ const char *e_ = getenv(%s); f77_assign(%s, e_ ? e_ : \"\", e_ ? strlen(e_) : 0); }" % (name[0].rsplit(", ", 1)[0], var[0])
        m = re.match(r"^CALL\s+(\w+)\s*\((.*)\)\s*$", txt, re.I | re.S)
        if m:
            callee = m.group(1).upper()
            if callee not in known_subs:
                # routine of another source file that is outside the hot path: calling it aborts
                u.stubs.add(callee)
                return "f77_untranslated(\"%s\");" % callee
            pre, cargs, hidden = [], [], []
            for a in split_top(m.group(2)):
                aup = a.strip().upper()
                if re.match(r"^\w+$", aup) and aup in u.vars and u.vars[aup]["type"] == "c":
                    cargs.append(u.cname(aup))
                    hidden.append(("len_" + u.cname(aup)) if aup in u.args else "(size_t)(%s)" % u.cstr(u.vars[aup]["clen"], "i"))
                elif re.match(r"^\w+$", aup) and aup in u.vars and aup not in u.params:
                    # (a PARAMETER constant that also has a type declaration is an expression: it goes through a temporary)
                    cargs.append(u.cname(aup) if (aup in u.args or u.vars[aup]["dims"]) else "&" + u.cname(aup))
                elif re.match(r"^\w+\s*\(.*\)$", aup) and aup.split("(")[0].strip() in u.vars and u.vars[aup.split("(")[0].strip()]["dims"]:
                    e = u.cexpr(a)
                    cargs.append("&" + e[0])
                else:
                    e = u.cexpr(a)
                    if e[1] == "c" and e[0].startswith('"'):            # a character constant: pointer + hidden length
                        lit, ln = e[0].rsplit(", ", 1)
                        cargs.append("(char *)" + lit)
                        hidden.append("(size_t)" + ln)
                        continue
                    if e[1] == "c":
                        raise Unsupported("character actual argument")
                    tmp = new_tmp()
                    pre.append("%s %s = %s;" % (CT[e[1]], tmp, e[0]))
                    cargs.append("&" + tmp)
            return "{ %s %s_(%s); }" % (" ".join(pre), callee.lower(), ", ".join(cargs + hidden))
        # assignment
        toks = tokenize(txt, u.defines)
        depth, eq = 0, None
        for i, (k, v) in enumerate(toks):
            if k == "op" and v == "(":
                depth += 1
            elif k == "op" and v == ")":
                depth -= 1
            elif k == "op" and v == "=" and depth == 0:
                eq = i
                break
        if eq is None:
            raise Unsupported("statement %r" % txt[:40])
        lhs = u.cexpr(toks[:eq])
        if any(k == "op" and v == "//" for k, v in toks[eq + 1:]):
            parts, cur, depth = [], [], 0
            for k, v in toks[eq + 1:]:
                if k == "op" and v == "(":
                    depth += 1
                if k == "op" and v == ")":
                    depth -= 1
                if k == "op" and v == "//" and depth == 0:
                    parts.append(cur)
                    cur = []
                else:
                    cur.append((k, v))
            parts.append(cur)
            code = "{ static char cat_[4096]; size_t n_ = 0; "
            for pt in parts:
                e = u.cexpr(pt)
                if e[1] != "c":
                    raise Unsupported("concatenation of a non-character operand")
                code += "n_ = f77_cat(cat_, n_, %s); " % e[0]
            return code + "f77_assign(%s, cat_, n_); }" % lhs[0]
        rhs = u.cexpr(toks[eq + 1:])
        if lhs[1] == "c" and rhs[1] == "c":
            return "f77_assign(%s, %s);" % (lhs[0], rhs[0])
        if lhs[1] == "c" or rhs[1] == "c":
            raise Unsupported("character assignment")
        return "%s = %s;" % (lhs[0], conv(rhs[0], rhs[1], lhs[1]))

    def close_do():
        ind.pop()
        emit("}")
        ind.pop()
        emit("}")

    for lab, txt, no in exe:
        up = txt.upper().strip()
        try:
            if up == "END":
                break
            if re.match(r"^FORMAT\s*\(", up):
                continue
            # terminal statement of labelled DO loops?
            terminal_for = [d for d in do_stack if d is not None and d == lab]
            m = re.match(r"^DO\s*WHILE\s*\((.*)\)\s*$", txt, re.I | re.S)
            if m:
                if lab:
                    emit("L%s: ;" % lab)
                emit("{")
                ind.append("  ")
                emit("while (%s) {" % u.cexpr(m.group(1))[0])
                ind.append("  ")
                do_stack.append(None)
                continue
            m = re.match(r"^DO\s+(?:(\d+)\s*,?\s*)?(\w+)\s*=\s*(.*)$", txt, re.I)
            if m and "=" in txt and not re.match(r"^DO\w*\s*=", up.replace(" ", "")[:0] or "x"):
                if lab:
                    emit("L%s: ;" % lab)
                parts = split_top(m.group(3))
                if len(parts) not in (2, 3):
                    raise Unsupported("DO bounds")
                var = m.group(2).upper()
                if var not in u.vars and not u.implicit_none:
                    u.vars[var] = {"type": u.implicit[var[0]], "dims": [], "clen": "1"}
                if var not in u.vars or u.vars[var]["type"] != "i":
                    raise Unsupported("DO variable %s" % var)
                e1, e2 = u.cstr(parts[0], "i"), u.cstr(parts[1], "i")
                e3 = u.cstr(parts[2], "i") if len(parts) == 3 else "1"
                v = u.scalar_ref(var)
                t = new_tmp()
                emit("{ int %s_s = %s; int %s_n; %s = %s; %s_n = ((%s) - %s + %s_s) / %s_s; if (%s_n < 0) %s_n = 0;" %
                     (t, e3, t, v, e1, t, e2, v, t, t, t, t))
                ind.append("  ")
                emit("for (; %s_n > 0; --%s_n, %s += %s_s) {" % (t, t, v, t))
                ind.append("  ")
                do_stack.append(m.group(1))
                continue
            if up in ("ENDDO", "END DO"):
                if not do_stack or do_stack[-1] is not None:
                    raise Unsupported("ENDDO without DO")
                if lab:
                    emit("L%s: ;" % lab)
                do_stack.pop()
                close_do()
                continue
            m = re.match(r"^IF\s*\((.*)\)\s*THEN$", txt, re.I | re.S)
            if m:
                if lab:
                    emit("L%s: ;" % lab)
                emit("if (%s) {" % u.cexpr(m.group(1))[0])
                ind.append("  ")
                continue
            m = re.match(r"^ELSE\s*IF\s*\((.*)\)\s*THEN$", txt, re.I | re.S)
            if m:
                ind.pop()
                emit("} else if (%s) {" % u.cexpr(m.group(1))[0])
                ind.append("  ")
                continue
            if up == "ELSE":
                ind.pop()
                emit("} else {")
                ind.append("  ")
                continue
            if up in ("ENDIF", "END IF"):
                ind.pop()
                emit("}")
                if lab:
                    emit("L%s: ;" % lab)
                continue
            if lab:
                emit("L%s: ;" % lab)
            if re.match(r"^IF\s*\(", up):                           # not IFIN = ... (a variable whose name starts with IF)
                # logical IF: find the matching parenthesis
                i = txt.index("(")
                depth, j = 0, i
                for j in range(i, len(txt)):
                    if txt[j] == "(":
                        depth += 1
                    elif txt[j] == ")":
                        depth -= 1
                        if depth == 0:
                            break
                cond, rest = txt[i + 1:j], txt[j + 1:].strip()
                if re.match(r"^\d+\s*,", rest):
                    raise Unsupported("arithmetic IF")
                emit("if (%s) { %s }" % (u.cexpr(cond)[0], simple_statement(rest)))
            else:
                emit(simple_statement(txt))
            # close every labelled DO that ends on this statement (innermost first)
            while do_stack and do_stack[-1] is not None and do_stack[-1] == lab:
                do_stack.pop()
                close_do()
        except Unsupported as e:
            raise Unsupported("%s line %d: %s" % (name, no, e))
    if do_stack:
        raise Unsupported("%s: unterminated DO" % name)
    # signature and declarations
    sig = []
    for a in args:
        v = u.vars[a]
        sig.append("%s *%s" % ("char" if v["type"] == "c" else CT[v["type"]], u.cname(a)))
    if has_char:
        sig += ["size_t len_%s" % u.cname(a) for a in args if u.vars[a]["type"] == "c"]
    decl = []
    for n, v in u.vars.items():
        if n in args:
            continue
        if v["type"] == "c":
            if len(v["dims"]) > 1:
                raise Unsupported("CHARACTER array %s of more than one dimension" % n)
            if v["dims"]:
                lo, hi = v["dims"][0]
                decl.append("  static char %s[((%s) - (%s) + 1) * (%s)];" % (u.cname(n), u.cstr(hi, "i"), u.cstr(lo, "i"), u.cstr(v["clen"], "i")))
                continue
            decl.append("  static char %s[%s];" % (u.cname(n), u.cstr(v["clen"], "i")))
            continue
        if v["dims"]:
            size = " * ".join("((%s) - (%s) + 1)" % (u.cstr(hi, "i"), u.cstr(lo, "i")) for lo, hi in v["dims"])
            if "(*" in size:                                   # automatic array sized by a dummy argument: stack, zeroed
                decl.append("  %s %s[%s]; memset(%s, 0, sizeof %s);" % (CT[v["type"]], u.cname(n), size, u.cname(n), u.cname(n)))
            else:
                decl.append("  static %s %s[%s];" % (CT[v["type"]], u.cname(n), size))
        else:
            decl.append("  %s %s = 0;" % (CT[v["type"]], u.cname(n)))
    # drop labels that no GOTO targets (avoids unused-label noise)
    body = [ln for ln in out if not (re.match(r"^\s*L(\d+): ;$", ln) and re.match(r"^\s*L(\d+): ;$", ln).group(1) not in u.labels_used)]
    code = ["void %s_(%s)" % (name.lower(), ", ".join(sig)), "{"] + decl + body + ["}", ""]
    return "\n".join(code), "void %s_(%s);" % (name.lower(), ", ".join(sig))


PRELUDE = r'''/* GENERATED by oracle/f77_to_c.py from the reference's Fortran sources -- do not commit (oracle/_ref/ is git-ignored) */
#include <math.h>
#include <stdlib.h>
#include <stddef.h>
#include <unistd.h>
/* libgcc's __powidf2 / __powisf2 (what gfortran emits for x**n with an integer n) */
static double f77_powi_d(double x, int m) { unsigned n = m < 0 ? -(unsigned)m : (unsigned)m; double y = (n % 2) ? x : 1; while (n >>= 1) { x = x * x; if (n % 2) y *= x; } return m < 0 ? 1 / y : y; }
static float f77_powi_f(float x, int m) { unsigned n = m < 0 ? -(unsigned)m : (unsigned)m; float y = (n % 2) ? x : 1; while (n >>= 1) { x = x * x; if (n % 2) y *= x; } return m < 0 ? 1 / y : y; }
/* minimal gfortran-compatible unformatted sequential I/O (4-byte record markers) */
#include <stdio.h>
#include <string.h>
static FILE *f77_fp[100]; static int f77_wr[100]; static unsigned char *f77_buf[100]; static size_t f77_len[100], f77_pos[100], f77_cap[100];
static int f77_index(const char *s, size_t ls, const char *t, size_t lt) { if (lt == 0 || lt > ls) return 0; for (size_t i = 0; i + lt <= ls; ++i) if (memcmp(s + i, t, lt) == 0) return (int)i + 1; return 0; }
static size_t f77_cat(char *buf, size_t n, const char *s, size_t ls) { if (n + ls > 4096) ls = 4096 - n; memcpy(buf + n, s, ls); return n + ls; }
static int f77_exists(const char *s, size_t ls) { char p[1024]; while (ls > 0 && s[ls - 1] == ' ') --ls; if (ls >= sizeof p) return 0; memcpy(p, s, ls); p[ls] = 0; FILE *f = fopen(p, "rb"); if (f) { fclose(f); return 1; } return 0; }
static void f77_system(const char *s, size_t ls) { char p[4200]; while (ls > 0 && s[ls - 1] == ' ') --ls; if (ls >= sizeof p) return; memcpy(p, s, ls); p[ls] = 0; if (system(p)) {} }
static void f77_assign(char *d, size_t ld, const char *s, size_t ls) { for (size_t i = 0; i < ld; ++i) d[i] = i < ls ? s[i] : ' '; }
static int f77_cmp(const char *a, size_t la, const char *b, size_t lb) { size_t n = la > lb ? la : lb; for (size_t i = 0; i < n; ++i) { char ca = i < la ? a[i] : ' ', cb = i < lb ? b[i] : ' '; if (ca != cb) return ca < cb ? -1 : 1; } return 0; }
/* list-directed READ from a text unit (or from a character string: pseudo-unit 100): a new record per statement, further records
   when the list needs them; items are numbers or unquoted / quoted character tokens */
static char f77_lline[101][4096]; static size_t f77_lpos[101];
static int f77_lbegin(int u) { if (u < 0 || u >= 100 || !f77_fp[u]) return 0; if (!fgets(f77_lline[u], 4096, f77_fp[u])) return -1; f77_lpos[u] = 0; return 1; }
static int f77_ibegin(const char *s, size_t ls) { if (ls > 4095) ls = 4095; memcpy(f77_lline[100], s, ls); f77_lline[100][ls] = 0; f77_lpos[100] = 0; return 1; }
static int f77_lsep(char c) { return c == ' ' || c == '\t' || c == ',' || c == '\n' || c == '\r'; }
static int f77_ltoken(int u, char *b, size_t cap) { for (;;) { char *p = f77_lline[u] + f77_lpos[u]; while (*p && f77_lsep(*p)) ++p;
  if (*p) { size_t n = 0; if (*p == '\'' || *p == '"') { char q = *p++; while (*p && *p != q && n + 1 < cap) b[n++] = *p++; if (*p == q) ++p; } else { while (*p && !f77_lsep(*p) && n + 1 < cap) b[n++] = *p++; }
    b[n] = 0; f77_lpos[u] = (size_t)(p - f77_lline[u]); return 1; }
  if (u == 100 || !f77_fp[u] || !fgets(f77_lline[u], 4096, f77_fp[u])) return 0; f77_lpos[u] = 0; } }
static int f77_lread(int u, double *v) { char b[128], *e; if (!f77_ltoken(u, b, sizeof b) || !b[0]) return 0; for (char *c = b; *c; ++c) if (*c == 'D' || *c == 'd') *c = 'e'; *v = strtod(b, &e); return *e == 0; }
static int f77_lread_c(int u, char *dst, size_t ld) { char b[4096]; if (!f77_ltoken(u, b, sizeof b)) return 0; size_t n = strlen(b); for (size_t i = 0; i < ld; ++i) dst[i] = i < n ? b[i] : ' '; return 1; }
static int f77_readline(int u, char *dst, size_t ld) { char b[4096]; if (u < 0 || u >= 100 || !f77_fp[u] || !fgets(b, sizeof b, f77_fp[u])) return 0; size_t n = strlen(b); while (n && (b[n - 1] == '\n' || b[n - 1] == '\r')) --n; f77_assign(dst, ld, b, n); return 1; }
static char f77_path[100][1024];
static int f77_open(int u, const char *name, size_t len, int status) { char *path = f77_path[u]; while (len > 0 && name[len - 1] == ' ') --len; if (len >= 1024) return 0; memcpy(path, name, len); path[len] = 0;
  if (f77_fp[u]) fclose(f77_fp[u]);
  f77_wr[u] = 0; status &= 3;
  if (status == 1) f77_fp[u] = fopen(path, "rb+"); else if (status == 2) { FILE *t = fopen(path, "rb"); if (t) { fclose(t); f77_fp[u] = 0; return 0; } f77_fp[u] = fopen(path, "wb+"); } else { f77_fp[u] = fopen(path, "rb+"); if (!f77_fp[u]) f77_fp[u] = fopen(path, "wb+"); }
  return f77_fp[u] != 0; }
static void f77_close(int u, int del) { if (u >= 0 && u < 100 && f77_fp[u]) { long p = ftell(f77_fp[u]); fflush(f77_fp[u]); if (p >= 0 && f77_wr[u]) { if (ftruncate(fileno(f77_fp[u]), p)) {} } fclose(f77_fp[u]); f77_fp[u] = 0; if (del) remove(f77_path[u]); } }
static void f77_rewind(int u) { if (f77_fp[u]) { fflush(f77_fp[u]); fseek(f77_fp[u], 0, SEEK_SET); } }
/* formatted records: the edit descriptors the reference uses on files (nX, Iw, Fw.d, Ew.d, Dw.d, A / Aw on output, Hollerith and quoted
   literals, groups with repeat counts) */
static struct { char kind; int w, d; const char *lit; } f77_ed[100][256]; static int f77_ned[100], f77_ied[100]; static char f77_line[100][4096]; static size_t f77_col[100]; static int f77_rd[100];
static const char *f77_fparse(int u, const char *p, int depth) { /* p points after '(' ; returns pointer after matching ')' */
  while (*p && *p != ')') { int rep = 0, has = 0; while (*p == ' ' || *p == ',') ++p; if (*p == ')') break; while (*p >= '0' && *p <= '9') { rep = rep * 10 + (*p - '0'); ++p; has = 1; } if (!has) rep = 1;
    if (*p == '(') { const char *q = p + 1, *e = q; for (int r = 0; r < rep; ++r) e = f77_fparse(u, q, depth + 1); p = e; continue; }
    if (*p == '\'') { const char *q = p + 1; while (*q && *q != '\'') ++q; if (f77_ned[u] < 256) { f77_ed[u][f77_ned[u]].kind = 'H'; f77_ed[u][f77_ned[u]].w = (int)(q - p - 1); f77_ed[u][f77_ned[u]].lit = p + 1; f77_ned[u]++; } p = *q ? q + 1 : q; continue; }
    char k = *p; if (k >= 'a' && k <= 'z') k -= 32; ++p;
    if (k == 'H') { /* Hollerith constant nHtext */ int n = 0; while (n < rep && p[n]) ++n; if (f77_ned[u] < 256) { f77_ed[u][f77_ned[u]].kind = 'H'; f77_ed[u][f77_ned[u]].w = n; f77_ed[u][f77_ned[u]].lit = p; f77_ned[u]++; } p += n; continue; }
    if (k == 'X') { if (f77_ned[u] < 256) { f77_ed[u][f77_ned[u]].kind = 'X'; f77_ed[u][f77_ned[u]].w = rep; f77_ned[u]++; } continue; }
    int w = 0, d = 0; while (*p >= '0' && *p <= '9') { w = w * 10 + (*p - '0'); ++p; } if (*p == '.') { ++p; while (*p >= '0' && *p <= '9') { d = d * 10 + (*p - '0'); ++p; } }
    for (int r = 0; r < rep && f77_ned[u] < 256; ++r) { f77_ed[u][f77_ned[u]].kind = k; f77_ed[u][f77_ned[u]].w = w; f77_ed[u][f77_ned[u]].d = d; f77_ned[u]++; } }
  return *p == ')' ? p + 1 : p; }
static int f77_fbegin(int u, const char *fmt, int rd) { if (!f77_fp[u]) return 0; f77_ned[u] = 0; f77_ied[u] = 0; f77_col[u] = 0; f77_rd[u] = rd; f77_fparse(u, fmt + 1, 0);
  if (rd) { if (!fgets(f77_line[u], sizeof f77_line[u], f77_fp[u])) return 0; size_t n = strlen(f77_line[u]); while (n && (f77_line[u][n - 1] == '\n' || f77_line[u][n - 1] == '\r')) f77_line[u][--n] = 0; } else f77_line[u][0] = 0;
  return 1; }
static int f77_fnext(int u) { for (;;) { if (f77_ied[u] >= f77_ned[u]) { /* format reversion: new record */ if (f77_rd[u]) { if (!fgets(f77_line[u], sizeof f77_line[u], f77_fp[u])) return -1; } else { fprintf(f77_fp[u], "%s\n", f77_line[u]); f77_wr[u] = 1; f77_line[u][0] = 0; } f77_col[u] = 0; f77_ied[u] = 0; }
    int i = f77_ied[u]++; if (f77_ed[u][i].kind == 'X') { if (f77_rd[u]) f77_col[u] += f77_ed[u][i].w; else { for (int k = 0; k < f77_ed[u][i].w; ++k) strcat(f77_line[u], " "); } continue; }
    if (f77_ed[u][i].kind == 'H') { if (f77_rd[u]) f77_col[u] += f77_ed[u][i].w; else strncat(f77_line[u], f77_ed[u][i].lit, (size_t)f77_ed[u][i].w); continue; } return i; } }
static void f77_fmt_e(char *out, double x, int w, int d, char ec) { char m[64], s[96]; if (x == 0.0) { snprintf(s, sizeof s, "0.%0*d%c+00", d, 0, ec); } else { snprintf(m, sizeof m, "%.*e", d - 1, fabs(x)); char *e = strchr(m, 'e'); int ex = atoi(e + 1) + 1; *e = 0; char dig[64]; int nd = 0; for (char *c = m; *c; ++c) if (*c != '.') dig[nd++] = *c; dig[nd] = 0;
    snprintf(s, sizeof s, "%s0.%s%c%c%02d", x < 0 ? "-" : "", dig, ec, ex >= 0 ? '+' : '-', abs(ex)); }
  size_t n = strlen(s); if ((int)n > w) { if (s[0] == '0') memmove(s, s + 1, n); else if (s[0] == '-' && s[1] == '0') memmove(s + 1, s + 2, n - 1); n = strlen(s); }
  if ((int)n > w) { memset(out, '*', w); out[w] = 0; } else snprintf(out, 96, "%*s", w, s); }
static void f77_fwrite(int u, double v, int is_int) { int i = f77_fnext(u); if (i < 0) return; char f[96]; char k = f77_ed[u][i].kind; int w = f77_ed[u][i].w, d = f77_ed[u][i].d;
  if (k == 'I') snprintf(f, sizeof f, "%*d", w, (int)v); else if (k == 'F') snprintf(f, sizeof f, "%*.*f", w, d, v); else f77_fmt_e(f, v, w, d, k == 'D' ? 'D' : 'E'); (void)is_int; strcat(f77_line[u], f); }
/* character item under an A / Aw descriptor: w absent = the length of the item; longer fields are padded on the left, shorter ones keep the first w characters */
static void f77_fwrite_a(int u, const char *s, size_t ls) { int i = f77_fnext(u); if (i < 0) return; size_t w = f77_ed[u][i].w > 0 ? (size_t)f77_ed[u][i].w : ls, n = strlen(f77_line[u]);
  if (n + w + 1 >= sizeof f77_line[u]) return; for (size_t k = ls; k < w; ++k) f77_line[u][n++] = ' '; for (size_t k = 0; k < (ls < w ? ls : w); ++k) f77_line[u][n++] = s[k]; f77_line[u][n] = 0; }
static int f77_fend(int u) { if (!f77_fp[u]) return 0;
  { /* literals that follow the last item, up to the next data descriptor */ int last = -1; for (int i = f77_ied[u]; i < f77_ned[u] && (f77_ed[u][i].kind == 'X' || f77_ed[u][i].kind == 'H'); ++i) if (f77_ed[u][i].kind == 'H') last = i;
    for (int i = f77_ied[u]; i <= last; ++i) { if (f77_ed[u][i].kind == 'X') { for (int k = 0; k < f77_ed[u][i].w; ++k) strcat(f77_line[u], " "); } else strncat(f77_line[u], f77_ed[u][i].lit, (size_t)f77_ed[u][i].w); } }
  fprintf(f77_fp[u], "%s\n", f77_line[u]); f77_wr[u] = 1; return 1; }
static int f77_fread(int u, double *v) { int i = f77_fnext(u); if (i < 0) return 0; int w = f77_ed[u][i].w, d = f77_ed[u][i].d; char f[128]; size_t n = strlen(f77_line[u]); int k = 0, dot = 0, expo = 0;
  for (int c = 0; c < w && k < 120; ++c) { char ch = (f77_col[u] + c < n) ? f77_line[u][f77_col[u] + c] : ' '; if (ch == ' ') continue; if (ch == 'D' || ch == 'd') ch = 'E'; if (ch == '.') dot = 1; if (ch == 'E' || ch == 'e') expo = 1; f[k++] = ch; } f[k] = 0; f77_col[u] += w;
  if (k == 0) { *v = 0.0; return 1; } char *end; *v = strtod(f, &end); if (end == f) return 0; if (!dot && !expo && f77_ed[u][i].kind != 'I') *v *= pow(10.0, -d); return 1; }
static void f77_untranslated(const char *name) { fprintf(stderr, "f77_to_c: call of untranslated routine %s\n", name); abort(); }
static int f77_eof(int u) { int c; if (!f77_fp[u]) return 1; c = fgetc(f77_fp[u]); if (c == EOF) return 1; ungetc(c, f77_fp[u]); return 0; }
static int f77_rbegin(int u) { int n = 0; if (!f77_fp[u] || fread(&n, 4, 1, f77_fp[u]) != 1 || n < 0) return 0; if ((size_t)n > f77_cap[u]) { f77_buf[u] = (unsigned char *)realloc(f77_buf[u], n); f77_cap[u] = n; }
  if (n && fread(f77_buf[u], 1, n, f77_fp[u]) != (size_t)n) return 0;
  int m = 0;
  if (fread(&m, 4, 1, f77_fp[u]) != 1 || m != n) return 0; f77_len[u] = n; f77_pos[u] = 0; return 1; }
static int f77_ritem(int u, void *dst, size_t sz) { if (f77_pos[u] + sz > f77_len[u]) return 0; memcpy(dst, f77_buf[u] + f77_pos[u], sz); f77_pos[u] += sz; return 1; }
static void f77_wbegin(int u) { f77_len[u] = 0; }
static void f77_witem(int u, const void *src, size_t sz) { if (f77_len[u] + sz > f77_cap[u]) { f77_cap[u] = 2 * (f77_len[u] + sz); f77_buf[u] = (unsigned char *)realloc(f77_buf[u], f77_cap[u]); } memcpy(f77_buf[u] + f77_len[u], src, sz); f77_len[u] += sz; }
static int f77_wend(int u) { int n = (int)f77_len[u]; if (!f77_fp[u]) return 0; f77_wr[u] = 1; return fwrite(&n, 4, 1, f77_fp[u]) == 1 && (n == 0 || fwrite(f77_buf[u], 1, n, f77_fp[u]) == (size_t)n) && fwrite(&n, 4, 1, f77_fp[u]) == 1; }
static int f77_powi_i(int x, int m) { int y = 1; if (m < 0) return (x == 1) ? 1 : ((x == -1) ? ((m % 2) ? -1 : 1) : 0); while (m-- > 0) y *= x; return y; }
'''


def subroutine_names(path):
    names = set()
    for lab, txt, no in logical_lines(path, {}):
        m = re.match(r"^SUBROUTINE\s+(\w+)", txt, re.I)
        if m:
            names.add(m.group(1).upper())
    return names


def translate_file(path, defines, wanted=None, known=None, skipped=None):
    """Returns (c_source, prototypes, report) for the SUBROUTINEs of one file."""
    defines = dict(defines)
    stmts = list(logical_lines(path, defines))
    units, cur = [], None
    for lab, txt, no in stmts:
        m = re.match(r"^SUBROUTINE\s+(\w+)\s*(?:\((.*)\))?\s*$", txt, re.I | re.S)
        if m:
            cur = [m.group(1).upper(), [a.strip().upper() for a in split_top(m.group(2) or "")], []]
            units.append(cur)
            continue
        if cur is not None:
            cur[2].append((lab, txt, no))
    names = {u[0] for u in units}
    src, protos, report = [], [], []
    ok = set()
    # two passes so that a routine may CALL one defined later in the file
    known = (set(names) | set(known or ())) - set(skipped or ())   # calls of `skipped` routines become run-time aborts
    for name, args, body in units:
        if wanted and name not in wanted:
            continue
        try:
            code, proto = translate_unit(name, args, body, defines, known)
            src.append(code)
            protos.append(proto)
            ok.add(name)
            report.append((name, "ok"))
        except Unsupported as e:
            report.append((name, "skipped: %s" % e))
    return "\n".join(src), protos, report


if __name__ == "__main__":
    defs = read_defines(sys.argv[2]) if len(sys.argv) > 2 else {}
    c, p, r = translate_file(sys.argv[1], defs)
    for name, st in r:
        print(name, st, file=sys.stderr)
    print(PRELUDE + "\n".join(p) + "\n\n" + c)
