"""Own reader for the CGS scene language (.geom files) -- SURVEY.md section 8(f) row N2.

A from-scratch interpreter for the dialect the reference accepts (reference src/cgs_read.cpp,
src/cgs_data.cpp, src/cgs.cpp:122-313, 510-575), producing the same document the reference parser
yields through oracle/_ref/scene_dump, so `Scene(parse_geom(...))` is interchangeable with the JSON
fixtures.  It mirrors the behaviours that decide results, including the odd ones:
  * statements end at newline or ';' (also inside comments); an opening ( [ { " captures through
    its matching close across lines, joined by blanks;
  * an expression splits at the FIRST operator of the loosest class present, classes from loosest
    to tightest: comparison, + -, ^, * /, then ?: and '.'  (so a-b-c == a-(b-c), 2*9/4*3 == 1.5);
  * undefined names evaluate to what strtod makes of them (usually 0), never an error;
  * range(a,b,step) has int((b-a)/step) elements; linspace is inclusive;
  * Composite(...) roots are collected newest first; lists become right-leaning binary chains;
    Complement flips `invert` for the FOLLOWING siblings only; Difference = Intersect[a, Complement[b..]];
  * Gaussian_source / CW_source / monitors / snapshot instances keep the reference's field names.
Validated field-for-field (exact floats) against the reference parser on every shipped scene that
parses (tests/test_cgs_parser.py).
"""
import math
import re

E_SUCCESS, E_BAD_SYNTAX, E_BAD_VALUE, E_BAD_TYPE, E_LACK_TOKENS, E_NAN, E_NOT_DEFINED, E_OUT_OF_RANGE = 0, 4, 5, 6, 2, 10, 11, 12
LIGHT_SPEED = 0.299792458
COMPONENT_IDS = {"Ex": 0, "Ey": 1, "Ez": 2, "Hx": 3, "Hy": 4, "Hz": 5}
DBL_MAX = 1.7976931348623157e308


class CgsError(Exception):
    def __init__(self, code, msg=""):
        super().__init__("CGS parse error %d %s" % (code, msg))
        self.code = code


class Vec3:
    __slots__ = ("el",)

    def __init__(self, x, y, z):
        self.el = [float(x), float(y), float(z)]


class Func:
    def __init__(self, name, min_args, impl):
        self.name, self.min_args, self.impl = name, min_args, impl


class Inst:
    """VAL_INST: an ordered stack of (name, value) pairs."""

    def __init__(self):
        self.fields = []

    def emplace(self, name, val):
        self.fields.append([name, val])

    def lookup(self, name):
        for k, v in reversed(self.fields):
            if k == name:
                return v
        return None

    def set_value(self, name, val):
        for f in reversed(self.fields):
            if f[0] == name:
                f[1] = val
                return
        self.emplace(name, val)

    @property
    def type(self):
        t = self.lookup("__type__")
        return t if isinstance(t, str) else None


def _is_num(v):
    return isinstance(v, float)


def _strtod(s):
    m = re.match(r"[ \t\n\r\f\v]*([+-]?(?:inf(?:inity)?|nan|(?:0[xX][0-9a-fA-F]*\.?[0-9a-fA-F]*(?:[pP][+-]?\d+)?)|(?:\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)))", s, re.I)
    if not m:
        return 0.0
    try:
        t = m.group(1)
        return float.fromhex(t) if re.match(r"[+-]?0[xX]", t) else float(t)
    except ValueError:
        return 0.0


def _cast_vec3(v):
    """value::cast_to(VAL_3VEC)."""
    if isinstance(v, Vec3):
        return Vec3(*v.el)
    if isinstance(v, list):
        if len(v) < 3:
            raise CgsError(E_LACK_TOKENS)
        if not all(_is_num(x) for x in v[:3]):
            raise CgsError(3)
        return Vec3(v[0], v[1], v[2])
    raise CgsError(E_BAD_VALUE)


# ------------------------------------------------------------------------------------------ text helpers
_OPEN = {"(": ")", "[": "]", "{": "}"}


def strchr_block(s, c):
    """first index of c outside any (), [], {}, quotes (cgs_read.cpp:38-71); -1 if none."""
    stk = []
    for i, ch in enumerate(s):
        if ch == c and not stk:
            return i
        if ch in "([{":
            stk.append(ch)
        elif ch in ")]}":
            if stk:
                stk.pop()
        elif ch == '"' or ch == "'":
            if stk and stk[-1] == ch:
                stk.pop()
            else:
                stk.append(ch)
    return -1


def _is_sep(c):
    return c in ("", " ", "\t", "\n", ";", "+", "-", "*", "/")


def token_block(s, comp):
    """first index of the token `comp` outside any block (cgs_read.cpp:76-118); -1 if none."""
    stk = []
    n = len(comp)
    for i, ch in enumerate(s):
        if not stk and s.startswith(comp, i):
            before = s[i - 1] if i > 0 else ""
            after = s[i + n] if i + n < len(s) else ""
            if (i == 0 or _is_sep(before)) and _is_sep(after):
                return i
        if ch in "([{":
            stk.append(ch)
        elif ch in ")]}":
            if stk:
                stk.pop()
        elif ch == '"' or ch == "'":
            if stk and stk[-1] == ch:
                stk.pop()
            else:
                stk.append(ch)
    return -1


def csv_to_list(s, sep=","):
    """split at top-level separators, dropping whitespace outside blocks/quotes (cgs_read.cpp:211-292)."""
    out, cur, stk = [], [], []
    verbatim = False
    i = 0
    while i < len(s):
        ch = s[i]
        if ch == sep and not stk:
            out.append("".join(cur))
            cur = []
        else:
            if ch == "\\":
                i += 1
                esc = s[i] if i < len(s) else ""
                rep = {"n": "\n", "t": "\t", "\\": "\\", '"': '"'}.get(esc)
                if rep is None:
                    raise CgsError(E_BAD_SYNTAX, "bad escape")
                cur.append(rep)
                i += 1
                continue
            elif ch == '"':
                if not stk or stk[-1] != '"':
                    stk.append('"')
                    verbatim = True
                else:
                    stk.pop()
                    verbatim = False
            elif ch in "([{":
                stk.append(ch)
            elif ch in ")]}":
                if not stk:
                    break
                if _OPEN.get(stk.pop()) != ch:
                    raise CgsError(E_BAD_SYNTAX, "mismatched block")
            if stk or verbatim or ch not in " \t\n":
                cur.append(ch)
        i += 1
    if stk:
        raise CgsError(E_BAD_SYNTAX, "unterminated block")
    # the trailing piece is kept whenever anything at all was written, even when it is empty itself
    # ("a,b," has three elements, "" has none)
    if out or cur:
        out.append("".join(cur))
    return out


# ------------------------------------------------------------------------------------------ interpreter
class Context(Inst):
    def __init__(self, parent=None):
        super().__init__()
        self.parent = parent

    def lookup(self, name):
        v = super().lookup(name)
        if v is None and self.parent is not None:
            return self.parent.lookup(name)
        return v

    # ---- expressions (context::parse_value, cgs_read.cpp:1780-2081) ----
    def parse_value(self, s):
        s = s.lstrip(" \n\t")
        is_numeric = bool(s) and (s[0] in "+-." or s[0].isdigit())
        first_open = last_close = -1
        stk = []
        nest = 0
        op_prec, op_loc = 0, 0
        n = len(s)
        i = 0
        while i < n:
            ch = s[i]
            if ch == "[":
                stk.append("["); nest += 1
                if first_open == -1:
                    first_open = i
            elif ch == "]":
                if not stk or stk.pop() != "[":
                    raise CgsError(E_BAD_SYNTAX, "unexpected ]")
                nest -= 1; last_close = i
            elif ch == "(":
                stk.append("("); nest += 1
                if first_open == -1:
                    first_open = i
            elif ch == ")":
                if not stk or stk.pop() != "(":
                    raise CgsError(E_BAD_SYNTAX, "unexpected )")
                nest -= 1; last_close = i
            elif ch == "{":
                stk.append("{"); nest += 1
                if first_open == -1:
                    first_open = i
            elif ch == "}":
                if not stk or stk.pop() != "{":
                    raise CgsError(E_BAD_SYNTAX, "unexpected }")
                nest -= 1; last_close = i
            elif ch == '"' and (i == 0 or s[i - 1] != "\\"):
                if not stk or stk[-1] != '"':
                    stk.append('"'); nest += 1
                    if first_open == -1:
                        first_open = i
                else:
                    stk.pop(); last_close = i; nest -= 1
            elif ch == "/":
                if i + 1 < n and s[i + 1] == "/":
                    break
            if nest == 0:
                nxt = s[i + 1] if i + 1 < n else ""
                if ((ch == "=" and nxt == "=") or ch == ">" or ch == "<") and op_prec < 5:
                    op_prec, op_loc = 5, i
                elif i != 0 and ch in "+-" and s[i - 1] != "e" and op_prec < 4:
                    op_prec, op_loc = 4, i
                elif ch == "^" and op_prec < 3:
                    op_prec, op_loc = 3, i
                elif ch in "*/" and op_prec < 2:
                    op_prec, op_loc = 2, i
                elif op_prec < 1 and (ch == "?" or (ch == "." and not is_numeric)):
                    op_prec, op_loc = 1, i
            i += 1
        if nest > 0:
            raise CgsError(E_BAD_SYNTAX, "expected close paren")
        if op_prec > 0:
            return self.do_op(s, op_loc)
        if first_open < 0 or last_close < 0:
            name = s.strip(" \t\n")
            v = self.lookup(name)
            if v is None:
                if name == "false":
                    return 0.0
                if name == "true":
                    return 1.0
                return _strtod(name)
            return v
        fo, lc = s[first_open], s[last_close]
        if fo == '"' and lc == '"':
            return s[first_open + 1:last_close]
        if fo == "[" and lc == "]":
            pre = s[:first_open]
            k = len(pre)
            while k > 0 and not _is_sep(pre[k - 1]):
                k -= 1
            pre_name = pre[k:]
            if pre_name == "":
                return self.parse_list(s[first_open:])
            lst = self.lookup(pre_name)
            if not isinstance(lst, list):
                raise CgsError(E_BAD_TYPE, "tried to index from non list type")
            idx = self.parse_value(s[first_open + 1:last_close])
            if not _is_num(idx):
                raise CgsError(E_BAD_TYPE, "only integers are valid indices")
            t = int(idx)
            if t < 0:
                t += len(lst)
            if t < 0 or t >= len(lst):
                raise CgsError(E_OUT_OF_RANGE)
            return lst[t]
        if fo == "{" and lc == "}":
            inst = Context(self)
            for el in csv_to_list(s[first_open + 1:last_close]):
                eq = strchr_block(el, "=")
                name, rval = (el[:eq].strip(" \t\n"), el[eq + 1:]) if eq >= 0 else (None, el)
                inst.emplace(name, inst.parse_value(rval))
            return inst
        if fo == "(" and lc == ")":
            if s[:first_open].strip(" \t\n") != "":
                return self.call(s[:last_close + 1], first_open)
            return self.parse_value(s[first_open + 1:last_close].strip(" \t\n"))
        # mismatched outer delimiters: the reference returns an undefined value
        return None

    def do_op(self, s, i):
        ch = s[i]
        width = 2 if (i + 1 < len(s) and s[i + 1] == "=") else 1
        left_s, right_s = s[:i], s[i + width:]
        if ch == "?":
            rest = s[i + 1:]
            col = strchr_block(rest, ":")
            if col < 0:
                raise CgsError(E_BAD_SYNTAX, "ternary without colon")
            cond = self.parse_value(left_s)
            if cond is None or (_is_num(cond) and cond == 0) or (not _is_num(cond) and False):
                return self.parse_value(rest[col + 1:])
            if not _is_num(cond):
                # non-numeric values have val.x read as raw union bits in the reference: treat as true
                pass
            return self.parse_value(s[i + width:i + 1 + col])
        if ch == ".":
            k = i
            while k > 0 and not _is_sep(s[k - 1]):
                k -= 1
            inst = self.lookup(s[k:i])
            if not isinstance(inst, Inst):
                raise CgsError(E_BAD_TYPE, "tried to lookup from non instance type")
            sub = Context(None)
            sub.fields = inst.fields
            sub.parent = getattr(inst, "parent", None)
            return sub.parse_value(s[i + 1:])
        lv = self.parse_value(left_s)
        rv = self.parse_value(right_s)
        if ch == "=":
            return 1.0 if _values_equal(lv, rv) else 0.0
        if ch in "><":
            if not (_is_num(lv) and _is_num(rv)):
                raise CgsError(E_BAD_VALUE)
            if width == 2:
                return 1.0 if (lv >= rv if ch == ">" else lv <= rv) else 0.0
            return 1.0 if (lv > rv if ch == ">" else lv < rv) else 0.0
        if ch == "+":
            if _is_num(lv) and _is_num(rv):
                return lv + rv
            if isinstance(lv, str) or isinstance(rv, str):
                return _rep_string(lv) + _rep_string(rv)
            return 0.0
        if ch == "-":
            return lv - rv if (_is_num(lv) and _is_num(rv)) else 0.0
        if ch == "*":
            return lv * rv if (_is_num(lv) and _is_num(rv)) else 0.0
        if ch == "/":
            if not _is_num(rv) or rv == 0:
                raise CgsError(E_NAN, "division by zero")
            return lv / rv if _is_num(lv) else 0.0
        if ch == "^":
            if not (_is_num(lv) and _is_num(rv)):
                raise CgsError(E_NAN)
            return math.pow(lv, rv)
        return 0.0

    # ---- lists (context::parse_list, cgs_read.cpp:1570-1641) ----
    def parse_list(self, s):
        start = s.find("[")
        end = strchr_block(s[start + 1:], "]")
        if start < 0 or end < 0:
            raise CgsError(E_BAD_SYNTAX)
        body = s[start + 1:start + 1 + end]
        f = token_block(body, "for")
        if f >= 0:
            rest = body[f + 3:]
            inn = token_block(rest, "in")
            if inn < 0:
                raise CgsError(E_BAD_SYNTAX)
            var = rest[:inn].strip(" \t\n")
            it = self.parse_value(rest[inn + 2:])
            if not isinstance(it, list):
                raise CgsError(E_BAD_VALUE)
            expr = body[:f]
            out = []
            self.emplace(var, None)
            for v in it:
                self.set_value(var, v)
                out.append(self.parse_value(expr))
            self.fields.pop()
            return out
        return [self.parse_value(el) for el in csv_to_list(body)]

    # ---- function calls (parse_func + dispatch, cgs_read.cpp:1431-1497, 1985-2066) ----
    def call(self, s, open_ind):
        name = s[:open_ind].strip(" \t\n")
        rest = s[open_ind + 1:]
        term = strchr_block(rest, ")")
        if term < 0:
            raise CgsError(E_BAD_SYNTAX)
        args, names = [], []
        for el in csv_to_list(rest[:term]):
            eq = strchr_block(el, "=")
            if eq >= 0:
                names.append(el[:eq]); el = el[eq + 1:]
            else:
                names.append(None)
            args.append(self.parse_value(el))
        if name in _BUILTINS:
            return _BUILTINS[name](args)
        if name in ("sin", "cos", "tan", "exp", "sqrt"):
            if len(args) < 1:
                raise CgsError(E_LACK_TOKENS)
            if not _is_num(args[0]):
                raise CgsError(E_BAD_TYPE, "math functions only accept numbers")
            return float(getattr(math, name)(args[0]))
        fv = self.lookup(name)
        if isinstance(fv, Func):
            if fv.min_args <= len(args):
                return fv.impl(self, args, names)
            return None
        raise CgsError(E_BAD_TYPE, "unrecognized function name %s" % name)

    # ---- statements (read_from_lines / read_single_line, cgs_read.cpp:2114-2262) ----
    def read_from_lines(self, lines):
        line, off = 0, 0
        while line < len(lines):
            if token_block(lines[line], "def") >= 0:
                raise CgsError(E_BAD_SYNTAX, "user-defined functions are not supported (the reference never runs them)")
            line, off = self._read_single(lines, line)
            # advance to the next line only when the statement consumed this one to its end
            if line < len(lines) and off >= len(lines[line]):
                line, off = line + 1, 0

    def _read_single(self, lines, line):
        """one statement starting at column 0 of lines[line]; returns the (line, off) where reading stopped.
        A block that closes on a later line ends the statement right there: whatever follows the closer on
        that line makes the reader start over from column 0 of it, as the reference does."""
        len0 = len(lines[line])
        buf = []                       # characters; "\0" marks the assignment splits
        off = 0
        started = False

        def get(l, o):
            return lines[l][o] if (l < len(lines) and o < len(lines[l])) else ""

        while True:
            ch = get(line, off)
            if off >= len0 or ch == "":
                break
            started = True
            if ch == "/" and off > 0 and get(line, off - 1) == "/":
                if len(buf) == 1:
                    started = False            # the line holds nothing but a comment
                else:
                    off = len(lines[line])
                buf.pop()
                break
            if ch == "*" and off > 0 and get(line, off - 1) == "/":
                if len(buf) == 1:
                    started = False
                buf.pop()
                # run forward (across lines) to the closing marker
                while True:
                    if off >= len(lines[line]):
                        if line == len(lines) - 1:
                            break
                        line, off = line + 1, 0
                    else:
                        off += 1
                    if get(line, off) == "*" and get(line, off + 1) == "/":
                        off += 2
                        break
                ch = get(line, off)
            if ch == "=":
                buf.append("\0")
                off += 1
                continue
            if ch and ch in "([{\"'":
                text, eline, eoff = _get_enclosed(lines, line, off, ch, _OPEN.get(ch, ch))
                if eline == line and eoff < len0:
                    buf.extend(text)
                    off = eoff
                    ch = get(line, off)
                else:
                    buf.extend(text + " ")
                    line, off = eline, eoff
                    break
            if ch:
                buf.append(ch)
            off += 1
        if started:
            text = "".join(buf)
            parts = text.split("\0")
            lval = parts[0].strip(" \t\n") if len(parts) > 1 else None
            self.emplace(lval, self.parse_value(parts[-1]))
            return line, off
        return line + 1, 0


def _get_enclosed(lines, li, off, start_delim, end_delim):
    """text from (li, off) through the matching end_delim with lines joined by ' ', and the position just
    after the closer -- or (len(lines), 0) when it never closes (get_enclosed + flatten(' '))."""
    depth = 0
    parts = []
    i, j = li, off
    while i < len(lines):
        line = lines[i]
        j0 = j
        ended = False
        while j < len(line):
            ch = line[j]
            if ch == start_delim:
                if end_delim == start_delim:
                    if j == 0 or line[j - 1] != "\\":
                        depth = 1 - depth
                        if depth == 0:
                            j += 1
                            ended = True
                            break
                else:
                    depth += 1
            elif ch == end_delim:
                depth -= 1
                if depth <= 0:
                    j += 1
                    ended = True
                    break
            j += 1
        parts.append(line[j0:j])
        if ended:
            return " ".join(parts), i, j
        i += 1
        j = 0
    return " ".join(parts), len(lines), 0


def _values_equal(a, b):
    if type(a) is not type(b):
        return False
    if _is_num(a) or isinstance(a, str):
        return a == b
    if isinstance(a, list):
        return len(a) == len(b) and all(_values_equal(x, y) for x, y in zip(a, b))
    return False


def _rep_string(v):
    if isinstance(v, str):
        return v
    if _is_num(v):
        return "%f" % v
    return ""


# ---- builtins (cgs_read.cpp:772-986) ----
def _vec(a):
    if len(a) < 3:
        raise CgsError(E_LACK_TOKENS)
    if not all(_is_num(x) for x in a[:3]):
        raise CgsError(3)
    return Vec3(a[0], a[1], a[2])


def _range(a):
    if len(a) == 0:
        raise CgsError(E_LACK_TOKENS)
    if not all(_is_num(x) for x in a[:3]):
        raise CgsError(E_BAD_TYPE)
    if len(a) == 1:
        if a[0] < 0:
            raise CgsError(E_BAD_VALUE)
        return [float(i) for i in range(int(a[0]))]
    if len(a) == 2:
        s0, e0 = int(a[0]), int(a[1])
        if e0 < s0:
            raise CgsError(E_BAD_VALUE)
        return [float(i) for i in range(s0, e0)]
    s0, e0, inc = a[0], a[1], a[2]
    if e0 < s0 or inc == 0:
        raise CgsError(E_BAD_VALUE)
    return [i * inc + s0 for i in range(int((e0 - s0) / inc))]


def _linspace(a):
    if len(a) < 3:
        raise CgsError(E_LACK_TOKENS)
    if not all(_is_num(x) for x in a[:3]):
        raise CgsError(E_BAD_TYPE)
    n = int(a[2])
    if n < 2:
        raise CgsError(E_BAD_VALUE)
    step = (a[1] - a[0]) / (n - 1)
    return [step * i + a[0] for i in range(n)]


def _flatten(a):
    if len(a) < 1:
        raise CgsError(E_LACK_TOKENS)
    if not isinstance(a[0], list):
        raise CgsError(E_BAD_VALUE)
    out = []

    def rec(l):
        for x in l:
            if isinstance(x, list):
                rec(x)
            else:
                out.append(x)
    rec(a[0])
    return out


def _print(a):
    return None


_BUILTINS = {"vec": _vec, "range": _range, "linspace": _linspace, "flatten": _flatten, "print": _print}


# ---- scene constructors (cgs_data.cpp:22-431) ----
def _named(args, names, key):
    for v, n in zip(args, names):
        if n == key:
            return v
    return None


def _mk(type_name):
    inst = Inst()
    inst.emplace("__type__", type_name)
    return inst


def _gaussian_source(ctx, a, names):
    if len(a) < 5:
        raise CgsError(E_LACK_TOKENS)
    if not isinstance(a[0], str) or not all(_is_num(x) for x in a[1:5]):
        raise CgsError(3)
    r = _mk("Gaussian_source")
    r.emplace("component", float(COMPONENT_IDS.get(a[0], 0)))
    r.emplace("wavelength", a[1]); r.emplace("amplitude", a[2]); r.emplace("width", a[3]); r.emplace("phase", a[4])
    r.emplace("cutoff", 5.0); r.emplace("start_time", 5.0)
    for v, n in list(zip(a, names))[5:]:
        if n == "cutoff":
            r.set_value("cutoff", v)
        elif n == "start_time":
            r.set_value("start_time", v)
    r.emplace("region", a[-1])
    return r


def _cw_source(ctx, a, names):
    if len(a) < 4:
        raise CgsError(E_LACK_TOKENS)
    if not isinstance(a[0], str) or not all(_is_num(x) for x in a[1:4]):
        raise CgsError(3)
    r = _mk("CW_source")
    r.emplace("component", float(COMPONENT_IDS.get(a[0], 0)))
    r.emplace("wavelength", a[1]); r.emplace("amplitude", a[2]); r.emplace("start_time", a[3])
    r.emplace("slowness", 5.0)
    r.emplace("end_time", a[4] if (len(a) > 4 and _is_num(a[4])) else DBL_MAX)
    for v, n in list(zip(a, names))[5:]:
        if n == "slowness":
            r.set_value("slowness", v)
    r.emplace("region", a[-1])
    return r


def _box(ctx, a, names):
    if len(a) < 2:
        raise CgsError(E_LACK_TOKENS)
    r = _mk("Box")
    r.emplace("pt_1", _cast_vec3(a[0])); r.emplace("pt_2", _cast_vec3(a[1]))
    return r


def _sub(a, b):
    return [a[i] - b[i] for i in range(3)]


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _norm(a):
    acc = 0.0
    for x in a:
        acc += x * x
    return math.sqrt(acc)


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _plane(ctx, a, names):
    if len(a) < 2:
        raise CgsError(E_LACK_TOKENS)
    if len(a) < 3:
        normal = _cast_vec3(a[0])
        if not _is_num(a[1]):
            raise CgsError(E_BAD_VALUE)
        offset = a[1]
    else:
        p1, p2, p3 = _cast_vec3(a[0]).el, _cast_vec3(a[1]).el, _cast_vec3(a[2]).el
        c = _cross(_sub(p2, p1), _sub(p3, p1))
        nn = _norm(c)
        normal = Vec3(c[0] / nn, c[1] / nn, c[2] / nn)
        offset = _dot(normal.el, p1)
    r = _mk("Plane")
    r.emplace("normal", normal); r.emplace("offset", offset)
    return r


def _sphere(ctx, a, names):
    if len(a) < 2:
        raise CgsError(E_LACK_TOKENS)
    cent = _cast_vec3(a[0])
    if not _is_num(a[1]):
        raise CgsError(E_BAD_VALUE)
    r = _mk("Sphere")
    r.emplace("center", cent); r.emplace("radius", a[1])
    return r


def _cylinder(ctx, a, names):
    if len(a) < 3:
        raise CgsError(E_LACK_TOKENS)
    cent = _cast_vec3(a[0])
    if not (_is_num(a[1]) and _is_num(a[2])):
        raise CgsError(E_BAD_VALUE)
    r2 = a[3] if (len(a) > 3 and _is_num(a[3])) else a[2]
    r = _mk("Cylinder")
    r.emplace("center", cent); r.emplace("h", a[1]); r.emplace("r1", a[2]); r.emplace("r2", r2)
    return r


def _as_list(v):
    """cast_to(VAL_LIST) as the constructors use it."""
    if isinstance(v, list):
        return v
    if isinstance(v, Vec3):
        return list(v.el)
    if isinstance(v, Inst):
        # brace form: the reference builds a list of the instance's size whose usable elements are not
        # geometry instances -> an object without children
        return [None] * len(v.fields)
    raise CgsError(E_BAD_VALUE)


def _composite(ctx, a, names):
    if len(a) < 1:
        raise CgsError(E_LACK_TOKENS)
    vl = _as_list(a[-1])
    r = _mk("Composite")
    for v, n in list(zip(a, names))[:-1]:
        if n:
            r.emplace(n, v)
    r.emplace("geometry", vl)
    return r


def _simple_group(type_name):
    def f(ctx, a, names):
        if len(a) < 1:
            raise CgsError(E_LACK_TOKENS)
        r = _mk(type_name)
        r.emplace("geometry", _as_list(a[0]))
        return r
    return f


def _difference(ctx, a, names):
    if len(a) < 1:
        raise CgsError(E_LACK_TOKENS)
    vl = _as_list(a[0])
    if len(vl) < 2:
        raise CgsError(E_BAD_VALUE)
    comp = _mk("Complement")
    comp.emplace("geometry", list(vl[1:]))
    r = _mk("Intersect")
    r.emplace("geometry", [vl[0], comp])
    return r


def _rotate(ctx, a, names):
    if len(a) < 3:
        raise CgsError(E_LACK_TOKENS)
    if not _is_num(a[0]):
        raise CgsError(E_BAD_TYPE)
    r = _mk("Rotate")
    r.emplace("axis", _cast_vec3(a[1])); r.emplace("theta", a[0]); r.emplace("geometry", _as_list(a[2]))
    return r


def _snapshot(ctx, a, names):
    if len(a) < 2:
        raise CgsError(E_LACK_TOKENS)
    if not isinstance(a[0], str):
        raise CgsError(E_BAD_VALUE)
    view = _cast_vec3(a[1])
    look = _named(a, names, "look")
    try:
        look = _cast_vec3(look) if look is not None else None
    except CgsError:
        look = None
    if look is None:
        look = Vec3(*[x * -1 for x in view.el])
    up = _named(a, names, "up")
    try:
        up = _cast_vec3(up) if up is not None else None
    except CgsError:
        up = None
    if up is None:
        up = Vec3(0, 0, 1)
    res = _named(a, names, "resolution")
    ns = _named(a, names, "n_samples")
    stp = _named(a, names, "step")
    sc = _named(a, names, "scale")
    r = _mk("snapshot")
    r.emplace("fname", a[0]); r.emplace("cam_v", view); r.emplace("look_v", look); r.emplace("up_v", up)
    r.emplace("res", res if _is_num(res) else 255.0)
    r.emplace("n_samples", ns if _is_num(ns) else 50000.0)
    r.emplace("step", stp if _is_num(stp) else 0.01)
    r.emplace("scale", sc if _is_num(sc) else 0.5 / _norm(look.el))
    return r


def _monitors(ctx, a, names):
    if len(a) < 1:
        raise CgsError(E_LACK_TOKENS)
    if not isinstance(a[0], list):
        raise CgsError(E_BAD_TYPE)
    r = _mk("monitor")
    r.emplace("locations", a[0])
    return r


def setup_geometry_context(con):
    for name, nmin, impl in (("Gaussian_source", 6, _gaussian_source), ("CW_source", 5, _cw_source), ("Box", 2, _box),
                             ("Plane", 2, _plane), ("Sphere", 2, _sphere), ("Cylinder", 2, _cylinder),
                             ("Composite", 1, _composite), ("Union", 1, _simple_group("Union")),
                             ("Intersect", 1, _simple_group("Intersect")), ("Complement", 1, _simple_group("Complement")),
                             ("Rotate", 1, _rotate), ("Difference", 1, _difference), ("snapshot", 2, _snapshot),
                             ("monitors", 1, _monitors)):
        con.emplace(name, Func(name, nmin, impl))


def context_from_settings(s):
    """reference src/disp.cpp:28-51."""
    con = Context()
    con.emplace("pi", math.pi)
    con.emplace("pml_thickness", float(s.pml_thickness))
    con.emplace("sim_length", float(s.len))
    con.emplace("length", 2 * s.pml_thickness + s.len)
    con.emplace("l_per_um", float(s.um_scale))
    con.emplace("out_dir", s.out_dir if s.out_dir is not None else "/tmp")

    def um_to_l(ctx, a, names):
        if len(a) < 1:
            raise CgsError(E_LACK_TOKENS)
        if not _is_num(a[0]):
            raise CgsError(3)
        sc = ctx.lookup("l_per_um")
        if not _is_num(sc):
            raise CgsError(E_NOT_DEFINED)
        return sc * a[0]

    def fs_to_t(ctx, a, names):
        if len(a) < 1:
            raise CgsError(E_LACK_TOKENS)
        if not _is_num(a[0]):
            raise CgsError(3)
        sc = ctx.lookup("l_per_um")
        if not _is_num(sc):
            raise CgsError(E_NOT_DEFINED)
        return LIGHT_SPEED * sc * a[0]
    con.emplace("um_to_l", Func("um_to_l", 1, um_to_l))
    con.emplace("fs_to_t", Func("fs_to_t", 1, fs_to_t))
    if s.user_opts:
        con.read_from_lines(split_lines(s.user_opts, only_semicolon=True))
    return con


def split_lines(text, only_semicolon=False):
    """line_buffer(fname): split at newline and ';', drop empty pieces; line_buffer(str, ';'): split at
    top-level ';' only (blocks and quotes protect it)."""
    if not only_semicolon:
        return [p for p in re.split(r"[;\n]", text) if len(p) > 0]
    out, cur, nest = [], [], 0
    i = 0
    while i < len(text):
        ch = text[i]
        if ch == '"':
            j = text.find('"', i + 1)
            j = len(text) if j < 0 else j
            cur.append(text[i:j + 1]); i = j + 1
            continue
        if ch in "([{":
            nest += 1
        elif ch in ")]}":
            nest -= 1
        if nest <= 0 and ch == ";":
            out.append("".join(cur)); cur = []
        else:
            cur.append(ch)
        i += 1
    out.append("".join(cur))
    return out


# ------------------------------------------------------------------------------------------ CSG trees
_IDENT = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]


def _matmul(a, b):
    out = [0.0] * 9
    for i in range(3):
        for j in range(3):
            acc = 0.0
            for k in range(3):
                acc += a[3 * i + k] * b[3 * k + j]
            out[3 * i + j] = acc
    return out


def make_rotation(theta, axis):
    """geometry.hpp:203-214."""
    n = _norm(axis)
    a = [x / n for x in axis]
    ct, st = math.cos(theta), math.sin(theta)
    ctc = 1.0 - ct
    r = [0.0] * 9
    r[0] = ct + a[0] * a[0] * ctc; r[1] = a[0] * a[1] * ctc - a[2] * st; r[2] = a[0] * a[2] * ctc + a[1] * st
    r[3] = r[1] + 2 * a[2] * st; r[4] = ct + a[1] * a[1] * ctc; r[5] = a[1] * a[2] * ctc - a[0] * st
    r[6] = r[2] - 2 * a[1] * st; r[7] = r[5] + 2 * a[0] * st; r[8] = ct + a[2] * a[2] * ctc
    return r


def _make_object(inst, invert, tstack):
    """make_object (cgs.cpp:122-199): returns (node | None, kind) with kind in
    'prim', 'composite', 'invert', 'rotate'."""
    t = inst.type
    if t == "Box":
        p1, p2 = _cast_vec3(inst.lookup("pt_1")).el, _cast_vec3(inst.lookup("pt_2")).el
        off = [(p2[i] - p1[i]) / 2 for i in range(3)]
        cen = [p1[i] + off[i] for i in range(3)]
        off = [o * -1 if o < 0 else o for o in off]
        return {"type": "box", "invert": invert, "center": cen, "offset": off}, "prim"
    if t == "Plane":
        off = inst.lookup("offset")
        if not _is_num(off):
            raise CgsError(E_BAD_TYPE)
        nv = _cast_vec3(inst.lookup("normal")).el
        nn = _norm(nv)
        return {"type": "plane", "invert": 0, "normal": [x / nn for x in nv], "offset": -off if invert else off}, "prim"
    if t == "Sphere":
        rad = inst.lookup("radius")
        if not _is_num(rad):
            raise CgsError(E_BAD_TYPE)
        return {"type": "sphere", "invert": invert, "center": _cast_vec3(inst.lookup("center")).el, "rad": rad}, "prim"
    if t == "Cylinder":
        h, r1, r2 = inst.lookup("h"), inst.lookup("r1"), inst.lookup("r2")
        if not (_is_num(h) and _is_num(r1) and _is_num(r2)):
            raise CgsError(E_BAD_TYPE)
        r1s = r1 * r1
        return {"type": "cylinder", "invert": invert, "center": _cast_vec3(inst.lookup("center")).el, "height": h,
                "r1_sq": r1s, "r1_sq_x_h": r1s * h, "r2_sq": r2 * r2}, "prim"
    if t in ("Union", "Intersect"):
        gv = inst.lookup("geometry")
        if not isinstance(gv, list):
            raise CgsError(E_BAD_TYPE)
        node = _new_composite(0 if t == "Union" else 1, invert)
        _init_from_list(node, gv, tstack)
        return node, "composite"
    if t == "Complement":
        return None, "invert"
    if t == "Rotate":
        return None, "rotate"
    raise CgsError(E_BAD_VALUE)


def _new_composite(cmb, invert=0):
    return {"type": "composite", "invert": invert, "cmb": cmb, "M": list(_IDENT), "children": [None, None]}


def _composite_from_inst(cmb, inst, invert, tstack):
    node = _new_composite(cmb, invert)
    geom = inst.fields[-1][1] if inst.fields else None
    if isinstance(geom, list):
        _init_from_list(node, geom, tstack)
    return node


def _init_from_list(node, lst, tstack):
    """composite_object::init_from_list (cgs.cpp:223-267)."""
    last = [node]
    cmb = node["cmb"]
    invert = 0
    trans = list(_IDENT)
    for m in reversed(tstack):
        trans = _matmul(trans, m)
    n = len(lst)

    def append(obj, is_last):
        lc = last[0]
        if lc["children"][0] is None:
            lc["children"][0] = obj
        elif is_last:
            lc["children"][1] = obj
        else:
            nc = _new_composite(cmb)
            lc["children"][1] = nc
            last[0] = nc
            nc["children"][0] = obj

    for i, g in enumerate(lst):
        if not isinstance(g, Inst):
            continue
        try:
            obj, kind = _make_object(g, invert, tstack)
        except CgsError:
            break
        if obj is None:
            if kind == "invert":
                invert = 1 - invert
                obj = _composite_from_inst(cmb, g, 1 - invert, tstack)
                append(obj, i == n - 1)
            elif kind == "rotate":
                va, vt = g.lookup("axis"), g.lookup("theta")
                if isinstance(va, Vec3) and _is_num(vt):
                    tstack.append(make_rotation(vt, va.el))
                    obj = _composite_from_inst(cmb, g, invert, tstack)
                    append(obj, i == n - 1)
        else:
            if kind == "prim":
                obj["M"] = list(trans)
            append(obj, i == n - 1)


# ------------------------------------------------------------------------------------------ document
def _to_doc(v):
    if isinstance(v, float):
        if math.isnan(v):
            return "nan"
        if math.isinf(v):
            return "inf" if v > 0 else "-inf"
        return v
    if isinstance(v, str):
        return v
    if isinstance(v, Vec3):
        return {"vec3": [_to_doc(x) for x in v.el]}
    if isinstance(v, list):
        return [_to_doc(x) for x in v]
    if isinstance(v, Func):
        return {"func": True}
    if isinstance(v, Inst):
        return {"__fields__": [[k if k is not None else "", _to_doc(x)] for k, x in v.fields]}
    return None


def _tree_doc(n):
    if n is None:
        return None
    d = {k: v for k, v in n.items() if k != "children"}
    if n["type"] == "composite":
        d["children"] = [_tree_doc(c) for c in n["children"]]
    if "M" not in d:
        d["M"] = list(_IDENT)
    return d


def parse_geom_text(text, settings):
    """scene::scene(fname, context) (cgs.cpp:510-575) -> the document Scene() consumes."""
    con = context_from_settings(settings)
    setup_geometry_context(con)
    init_size = len(con.fields)
    ercode = E_SUCCESS
    try:
        con.read_from_lines(split_lines(text))
    except CgsError as e:
        ercode = e.code
    roots = []
    if ercode == E_SUCCESS:
        for name, inst in reversed(con.fields[init_size:]):          # newest first
            if isinstance(inst, Inst) and inst.type == "Composite":
                ct = inst.lookup("combine_type")
                cmb = 1 if (isinstance(ct, str) and ct in ("intersect", "Intersect")) else 0
                tree = _composite_from_inst(cmb, inst, 0, [])
                meta = {k: _to_doc(v) for k, v in inst.fields[1:-1]}
                roots.append({"thickness": 1.0, "metadata": meta, "tree": _tree_doc(tree)})
    return {"ercode": ercode, "n_roots": len(roots), "roots": roots, "context": _to_doc(con)}


def parse_geom(path, settings):
    with open(path, "r") as fp:
        return parse_geom_text(fp.read(), settings)
