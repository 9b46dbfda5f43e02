"""Regular-expression surface syntax of Kleenex `/.../` terms.

Restates the dialect the reference gets from `fancyRegexParser` with `/`
illegal and free-spacing off (src/KMC/Kleenex/Parser.hs:204-206;
regexps-syntax/KMC/Syntax/Parser.hs:55-242, Config.hs:44-58).  The AST is a
tagged tuple mirroring regexps-syntax/KMC/Syntax/External.hs:36-56:

  ("one",) ("dot",) ("chr", codepoint) ("group", re) ("concat", a, b)
  ("branch", a, b) ("class", positive, [(lo, hi)]) ("range", re, n, m|None)
  ("star", re) ("lazystar", re) ("plus", re) ("lazyplus", re)
  ("question", re) ("lazyquestion", re) ("suppress", re)
"""


class RegexSyntaxError(Exception):
    pass


_ESCAPES = {"n": "\n", "t": "\t", "r": "\r", "a": "\a", "f": "\f", "v": "\v"}
_HEX = "0123456789abcdefABCDEF"


class _P:
    def __init__(self, s, i, illegal):
        self.s = s
        self.i = i
        self.illegal = illegal

    def peek(self, k=0):
        j = self.i + k
        return self.s[j] if j < len(self.s) else ""

    def startswith(self, t):
        return self.s.startswith(t, self.i)

    def fail(self, msg):
        raise RegexSyntaxError("%s at offset %d" % (msg, self.i))

    # legalChar (Parser.hs:147-174).  mode: "no" | "first" | "in"
    def legal_char(self, mode):
        if mode == "no":
            notchars = "*|()\\" + self.illegal + ".$^[?+{"
        elif mode == "in":
            notchars = "]"
        else:
            notchars = ""
        c = self.peek()
        if c == "":
            return None
        if c == "\\":
            d = self.peek(1)
            if d and (d in _ESCAPES or d in notchars):
                self.i += 2
                return _ESCAPES.get(d, d)
            if d == "x":
                # \xFF or \x{F...}
                if self.peek(2) in _HEX and self.peek(2) != "" and self.peek(3) in _HEX and self.peek(3) != "":
                    v = int(self.s[self.i + 2:self.i + 4], 16)
                    self.i += 4
                    return chr(v)
                if self.peek(2) == "{":
                    j = self.i + 3
                    k = j
                    while k < len(self.s) and self.s[k] in _HEX:
                        k += 1
                    if k > j and k < len(self.s) and self.s[k] == "}":
                        self.i = k + 1
                        return chr(int(self.s[j:k], 16))
            if d == "u":
                h = self.s[self.i + 2:self.i + 6]
                if len(h) == 4 and all(ch in _HEX for ch in h):
                    self.i += 6
                    return chr(int(h, 16))
            # fall through: a bare backslash is only legal where it is not special
        if c in notchars:
            return None
        self.i += 1
        return c

    def number(self, at_least_one=False):
        j = self.i
        while self.peek() != "" and self.peek() in "0123456789":
            self.i += 1
        if self.i == j:
            if at_least_one:
                return None
            return 0
        return int(self.s[j:self.i])

    def klass(self):
        # classP (Parser.hs:199-206); caller consumed "["
        positive = True
        if self.peek() == "^":
            positive = False
            self.i += 1
        ranges = []
        first = True
        while True:
            c1 = self.legal_char("first" if first else "in")
            if c1 is None:
                if first:
                    self.fail("empty character class")
                break
            first = False
            save = self.i
            if self.peek() == "-":
                self.i += 1
                c2 = self.legal_char("in")
                if c2 is None:
                    self.i = save
                    ranges.append((ord(c1), ord(c1)))
                else:
                    ranges.append((ord(c1), ord(c2)))
            else:
                ranges.append((ord(c1), ord(c1)))
        if self.peek() != "]":
            self.fail("expected ]")
        self.i += 1
        return ("class", positive, ranges)

    def atom(self):
        if self.startswith("(?:"):
            self.i += 3
            e = self.regex()
            if self.peek() != ")":
                self.fail("expected )")
            self.i += 1
            return ("group", e)
        if self.peek() == "(":
            self.i += 1
            e = self.regex()
            if self.peek() != ")":
                self.fail("expected )")
            self.i += 1
            return ("group", e)
        if self.startswith("[[:"):
            self.fail("POSIX named sets are not supported (Desugaring.hs:121)")
        if self.peek() == "[":
            self.i += 1
            return self.klass()
        if self.startswith("$("):
            self.i += 2
            e = self.regex()
            if not self.startswith(")$"):
                self.fail("expected )$")
            self.i += 2
            return ("suppress", e)
        if self.peek() == ".":
            self.i += 1
            return ("dot",)
        c = self.legal_char("no")
        if c is None:
            return None
        return ("chr", ord(c))

    def postfix(self, e):
        # one postfix operator per term (Parser.hs:99-125)
        if self.startswith("*?"):
            self.i += 2
            return ("lazystar", e)
        if self.peek() == "*":
            self.i += 1
            return ("star", e)
        if self.startswith("??"):
            self.i += 2
            return ("lazyquestion", e)
        if self.peek() == "?":
            self.i += 1
            return ("question", e)
        if self.startswith("+?"):
            self.i += 2
            return ("lazyplus", e)
        if self.peek() == "+":
            self.i += 1
            return ("plus", e)
        if self.peek() == "{":
            self.i += 1
            n = self.number()
            if self.peek() == ",":
                self.i += 1
                m = self.number(at_least_one=True)
            else:
                m = n
            if self.peek() != "}":
                self.fail("malformed range")
            self.i += 1
            if self.peek() == "?":
                self.fail("lazy ranges are not supported (Desugaring.hs:122)")
            return ("range", e, n, m)
        return e

    def factor(self):
        a = self.atom()
        if a is None:
            return None
        return self.postfix(a)

    def concat(self):
        fs = []
        while True:
            if fs and self.peek() == "|":
                break
            f = self.factor()
            if f is None:
                break
            fs.append(f)
        if not fs:
            self.fail("expected regular expression")
        e = fs[-1]
        for f in reversed(fs[:-1]):
            e = ("concat", f, e)
        return e

    def regex(self):
        a = self.concat()
        if self.peek() == "|":
            self.i += 1
            return ("branch", a, self.regex())
        return a


def parse_regex_at(s, i, illegal="/"):
    """Parse an (optionally anchored) regex starting at s[i]; returns
    (ast, next_index).  Anchors are recognised and dropped, as the Kleenex
    parser takes `snd` of the anchored result (src/KMC/Kleenex/Parser.hs:205)."""
    p = _P(s, i, illegal)
    if p.peek() == "^":
        p.i += 1
    e = p.regex()
    if p.peek() == "$":
        p.i += 1
    return e, p.i


def parse_regex(s):
    e, j = parse_regex_at(s, 0, illegal="")
    if j != len(s):
        raise RegexSyntaxError("unexpected %r at offset %d" % (s[j], j))
    return e
