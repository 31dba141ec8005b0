"""Host-side API wrappers of the reference's `Regex` that are pure plumbing over the batch results of
the scan path (reference regex.go:307-1610): first-match forms, text forms,
ReplaceAll*/Expand, Split, iterators.

The reference runs its own per-match loop in each of them (FindIndicesAt / FindSubmatchAt from the
previous end, with the rule that an empty match right behind a non-empty one is skipped —
regex.go:797-840, :1046-1104, :1143-1188, :1487-1515).  That loop is exactly the FindAll loop
(meta/findall.go:221-290), so here every wrapper takes ONE batch call — FindAllIndex or
FindAllSubmatchIndex on the device — and stitches the answer on the host.

The functions in the first half are pure (match list in, result out) and are tested on the CPU
against the reference's own vectors with the oracle's match lists; `RegexWrappers` binds them to
the device-backed `Regex`.  Text forms take and return `str`; offsets are BYTE offsets into the
UTF-8 encoding, as in Go."""

_SPECIAL = b"\\.+*?()|[]{}^$"


def _b(s):
    return s.encode("utf-8", "surrogateescape") if isinstance(s, str) else bytes(s)


def _s(b):
    return bytes(b).decode("utf-8", "surrogateescape")


def quote_meta(s):
    """reference regex.go:233 QuoteMeta (byte-wise, the 15 bytes of `special`)"""
    text = isinstance(s, str)
    out = bytearray()
    for c in _b(s):
        if c in _SPECIAL:
            out.append(0x5C)
        out.append(c)
    return _s(out) if text else bytes(out)


def expand(dst, template, src, match):
    """reference regex.go:951 expand: $0-$9 (ONE digit), `$$` -> `$`, everything else literal —
    `${name}` is not expanded by the reference (:977-981) and `$10` is group 1 followed by `0`.
    match: flat offsets (2 per group, -1 = unset).  Appends to the bytearray dst and returns it."""
    t, i = template, 0
    while i < len(t):
        if t[i] != 0x24 or i + 1 >= len(t):
            dst.append(t[i])
            i += 1
            continue
        nxt = t[i + 1]
        if 0x30 <= nxt <= 0x39:
            g = 2 * (nxt - 0x30)
            if g + 1 < len(match) and match[g] >= 0:
                dst += src[match[g]:match[g + 1]]
            i += 2
        elif nxt == 0x24:
            dst.append(0x24)
            i += 2
        else:  # `${` and unknown escapes: the `$` is literal
            dst.append(0x24)
            i += 1
    return dst


def replace_all_literal(src, pairs, repl):
    """reference regex.go:790 ReplaceAllLiteral over the FindAllIndex list of src"""
    out, last = bytearray(), 0
    for s, e in pairs:
        out += src[last:s]
        out += repl
        last = e
    out += src[last:]
    return bytes(out)


def replace_all(src, rows, repl):
    """reference regex.go:1006 ReplaceAll over the FindAllSubmatchIndex rows of src"""
    out, last = bytearray(), 0
    for m in rows:
        out += src[last:m[0]]
        expand(out, repl, src, m)
        last = m[1]
    out += src[last:]
    return bytes(out)


def replace_all_func(src, pairs, fn):
    """reference regex.go:1136 ReplaceAllFunc"""
    out, last = bytearray(), 0
    for s, e in pairs:
        out += src[last:s]
        out += fn(src[s:e])
        last = e
    out += src[last:]
    return bytes(out)


def split(s, pairs, n):
    """reference regex.go:1288 Split over the FindAllIndex list of s (bytes in, list of bytes out)"""
    if n == 0:
        return None
    if not pairs:
        return [s]
    out, last = [], 0
    for a, b in pairs:
        if last == 0 and a == 0 and b == 0:   # empty match at the very beginning (:1312-1316)
            continue
        if a == len(s) and b == len(s):       # empty match at the very end (:1318-1321)
            break
        out.append(s[last:a])
        last = b
        if n > 0 and len(out) >= n - 1:
            out.append(s[last:])
            return out
    out.append(s[last:])
    return out


def groups_of(src, row, text=False):
    """one submatch row -> list of group texts, None for an unset group (regex.go:621-690)"""
    conv = _s if text else bytes
    return [conv(src[row[k]:row[k + 1]]) if row[k] >= 0 and row[k + 1] >= 0 else (None if not text else "")
            for k in range(0, len(row), 2)]


class RegexWrappers:
    """Mixed into coregex_b200.Regex: needs FindAllIndex, FindAllSubmatchIndex, Match, Count,
    NumSubexp, String, _subexp_name."""

    # -- first match (reference regex.go:307-372, :621-690): the first element of the FindAll list
    def FindIndex(self, b):
        m = self.FindAllIndex(b, 1)
        return m[0] if m else None

    def Find(self, b):
        b = _b(b)
        m = self.FindIndex(b)
        return b[m[0]:m[1]] if m else None

    def FindStringIndex(self, s):
        return self.FindIndex(_b(s))

    def FindString(self, s):
        m = self.Find(_b(s))
        return _s(m) if m is not None else ""

    def FindSubmatchIndex(self, b):
        m = self.FindAllSubmatchIndex(b, 1)
        return m[0] if m else None

    def FindSubmatch(self, b):
        b = _b(b)
        m = self.FindSubmatchIndex(b)
        return groups_of(b, m) if m else None

    def FindStringSubmatchIndex(self, s):
        return self.FindSubmatchIndex(_b(s))

    def FindStringSubmatch(self, s):
        b = _b(s)
        m = self.FindSubmatchIndex(b)
        return groups_of(b, m, text=True) if m else None

    # -- all matches as text (reference regex.go:376-440, :764-788, :1376-1480)
    def FindAll(self, b, n=-1):
        b = _b(b)
        m = self.FindAllIndex(b, n)
        return [b[s:e] for s, e in m] if m else None

    def FindAllString(self, s, n=-1):
        m = self.FindAll(_b(s), n)
        return [_s(x) for x in m] if m else None

    def FindAllStringIndex(self, s, n=-1):
        return self.FindAllIndex(_b(s), n)

    def AppendAllIndex(self, dst, b, n=-1):
        """reference regex.go:748: appends (start, end) tuples to the caller's list"""
        dst.extend((s, e) for s, e in (self.FindAllIndex(b, n) or []))
        return dst

    def AppendAllStringIndex(self, dst, s, n=-1):
        return self.AppendAllIndex(dst, _b(s), n)

    def FindAllSubmatch(self, b, n=-1):
        b = _b(b)
        rows = self.FindAllSubmatchIndex(b, n)
        return [groups_of(b, r) for r in rows] if rows else None

    def FindAllStringSubmatch(self, s, n=-1):
        b = _b(s)
        rows = self.FindAllSubmatchIndex(b, n)
        return [groups_of(b, r, text=True) for r in rows] if rows else None

    def FindAllStringSubmatchIndex(self, s, n=-1):
        return self.FindAllSubmatchIndex(_b(s), n)

    def MatchString(self, s):
        return self.Match(_b(s))

    def CountString(self, s, n=-1):
        return self.Count(_b(s), n)

    # -- iterators (reference regex.go:1485-1580): the batch list, yielded
    def AllIndex(self, b):
        for s, e in self.FindAllIndex(b, -1) or []:
            yield (s, e)

    def AllStringIndex(self, s):
        return self.AllIndex(_b(s))

    def All(self, b):
        b = _b(b)
        for s, e in self.AllIndex(b):
            yield b[s:e]

    def AllString(self, s):
        for m in self.All(_b(s)):
            yield _s(m)

    # -- replace / expand / split (reference regex.go:790-1348)
    def ReplaceAllLiteral(self, src, repl):
        src = _b(src)
        return replace_all_literal(src, self.FindAllIndex(src, -1) or [], _b(repl))

    def ReplaceAllLiteralString(self, src, repl):
        return _s(self.ReplaceAllLiteral(_b(src), _b(repl)))

    def Expand(self, dst, template, src, match):
        return expand(dst, _b(template), _b(src), match)

    def ExpandString(self, dst, template, src, match):
        return expand(dst, _b(template), _b(src), match)

    def ReplaceAll(self, src, repl):
        src, repl = _b(src), _b(repl)
        if b"$" not in repl:  # :1017-1020
            return replace_all_literal(src, self.FindAllIndex(src, -1) or [], repl)
        return replace_all(src, self.FindAllSubmatchIndex(src, -1) or [], repl)

    def ReplaceAllString(self, src, repl):
        return _s(self.ReplaceAll(_b(src), _b(repl)))

    def ReplaceAllFunc(self, src, fn):
        src = _b(src)
        return replace_all_func(src, self.FindAllIndex(src, -1) or [], fn)

    def ReplaceAllStringFunc(self, src, fn):
        b = _b(src)
        return _s(replace_all_func(b, self.FindAllIndex(b, -1) or [], lambda m: _b(fn(_s(m)))))

    def Split(self, s, n=-1):
        b = _b(s)
        parts = split(b, self.FindAllIndex(b, -1) or [], n)
        return None if parts is None else [_s(p) for p in parts]

    # -- readers (reference regex.go:1619-1670: the reference drains the RuneReader into a string and
    # calls the string form; a Python text stream's read() is that drain)
    @staticmethod
    def _drain(reader):
        data = reader.read()
        return data if isinstance(data, str) else _s(data)

    def MatchReader(self, reader):
        return self.MatchString(self._drain(reader))

    def FindReaderIndex(self, reader):
        return self.FindStringIndex(self._drain(reader))

    def FindReaderSubmatchIndex(self, reader):
        return self.FindStringSubmatchIndex(self._drain(reader))

    # -- group names, copies, text marshalling (reference regex.go:575-604, :1585-1616)
    def SubexpNames(self):
        return [self._subexp_name(i) for i in range(self.NumSubexp() + 1)]

    def SubexpIndex(self, name):
        if name == "":
            return -1
        names = self.SubexpNames()
        return names.index(name) if name in names else -1

    def Copy(self):
        re = type(self)(self.String())
        if getattr(self, "_longest", False):
            re.Longest()
        return re

    def MarshalText(self):
        return self.String().encode()
