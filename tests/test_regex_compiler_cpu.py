"""CPU tier: the product's regex compiler (tables) interpreted by the test-only emulation library
(tests/emul) must agree with the oracle (PCRE2) and with Python `re` -- two independent
Perl-semantics engines -- on golden vectors and on seeded random patterns/inputs."""
import json
import os
import random
import re

import pytest

from oracle import oracle as orc
from tests.emul.emul import EmulRegex

HERE = os.path.dirname(os.path.abspath(__file__))


def _misc():
    with open(os.path.join(HERE, "golden", "ref_misc.json"), encoding="utf-8") as f:
        return json.load(f)


def test_golden_vectors():
    d = _misc()
    for c in d["prefix_search"] + d["multiline_start"]:
        r = EmulRegex(c["pattern"])
        assert r.supported, r.error
        assert [r.prefix_match(i.encode()) for i in c["inputs"]] == c["expected"], c["source"]
    for c in d["full_match_fields"]:
        r = EmulRegex(c["pattern"])
        assert r.supported, r.error
        b = c["input"].encode()
        caps = r.full_match(b)
        assert caps is not None and [b[o:o + l].decode() for o, l in caps] == c["fields"], c["source"]


UNIT_PATTERNS = [r"Exception.*", r"\s+at\s.*", r"\s*\.\.\.\d+ more", r"(\w+)\t(\w+).*"]


def test_multiline_unit_patterns_prefix():
    lines = [b"Exception in thread 'main' java.lang.NullPointerException",
             b"    at com.example.myproject.Book.getTitle(Book.java:16)", b"    ...23 more", b"unmatch log", b"",
             b" at x", b"...1 more"]
    for p in UNIT_PATTERNS:
        e, o = EmulRegex(p), orc.Regex(p)
        assert e.supported, e.error
        for ln in lines:
            assert e.prefix_match(ln) == o.prefix_match(ln), (p, ln)


@pytest.mark.parametrize("pattern,why", [
    (r"(a)\1", "back-reference"), (r"a(?=bc)", "look-ahead"), (r"a(?!b|cd)", "look-ahead"), (r"(?<=a)b", "look-behind"),
    (r"(a*)*", "empty"),
    (r"a++", "possessive"), (r"(?>a)", "atomic"), (r"\Zx", "escape"), (r"\Gx", "escape"),
])
def test_unsupported_is_reported_not_guessed(pattern, why):
    r = EmulRegex(pattern)
    assert not r.supported
    assert why in r.error


def test_not_word_boundary_inside_the_value_agrees_with_pcre2_and_python():
    """\\B (boost: match_within_word).  Away from the edges of the value boost, Perl, PCRE2 and Python agree: both
    neighbours are word characters or both are not.  AT an edge boost alone says false (perl_matcher_common.hpp:
    position == last / position == backstop), so a pattern that needs \\B there does not match here either."""
    rng = random.Random(77)
    pats = [r"(\w+)\B(\w)(.*)", r"(.*?)a\Bb(.*)", r"([a-z ]*) \B-(\S*)", r"(\w)\B(\w*)\b(.*)", r"x(?:\B|,)y(.*)"]
    alpha = "ab x-,y_1 "
    for p in pats:
        e, o, py = EmulRegex(p), orc.Regex(p), re.compile(p.encode(), re.S | re.M)
        assert e.supported, (p, e.error)
        for _ in range(400):
            v = "".join(rng.choice(alpha) for _ in range(rng.randint(0, 12))).encode()
            got, want = e.full_match(v), o.full_match(v)
            m = py.fullmatch(v)
            assert (got is None) == (want is None) == (m is None), (p, v)
            if got is not None:
                assert got == want, (p, v)
    # the boost-only rule at the edges of the value
    edge = EmulRegex(r"\B-(.*)")
    assert edge.supported and edge.full_match(b"-x") is None        # Perl / PCRE2 would match (start, next to '-')
    assert orc.Regex(r"\B-(.*)").full_match(b"-x") is not None
    tail = EmulRegex(r"(.*)-\B")
    assert tail.full_match(b"x-") is None                           # Perl / PCRE2 would match (end, after '-')
    assert EmulRegex(r"(.*)a\Bb").full_match(b"xab") is not None


def test_word_start_and_end_assertions():
    """\\< and \\> (boost: match_word_start / match_word_end) = \\b(?=\\w) and \\b(?<=\\w); the oracle hands PCRE2 that
    translation (PCRE2 reads \\< as a literal '<'), Python gets the same one as an independent check."""
    rng = random.Random(99)
    pats = [r"(.*?)\<(\w+)\>(.*)", r"\<(\w+) (.*)", r"(.*) (\w+)\>", r"(\W*)\<a(.*)", r"(.*)b\>(\W*)", r"(a|\<b)(.*)",
            r"(\w*)\>(.?)(.*)"]
    alpha = "ab_1 -,."
    for p in pats:
        e, o = EmulRegex(p), orc.Regex(p)
        py = re.compile(p.replace(r"\<", r"\b(?=\w)").replace(r"\>", r"\b(?<=\w)").encode(), re.S | re.M)
        assert e.supported, (p, e.error)
        for _ in range(500):
            v = "".join(rng.choice(alpha) for _ in range(rng.randint(0, 10))).encode()
            got, want = e.full_match(v), o.full_match(v)
            m = py.fullmatch(v)
            assert (got is None) == (want is None) == (m is None), (p, v, got, want)
            if got is not None:
                assert got == want, (p, v)
            if e.tdfa_info["states"] > 1:
                assert e.full_match_tdfa(v, 0) == want and e.full_match_tdfa(v, 1) == want, (p, v)
            assert e.prefix_match(v) == o.prefix_match(v), (p, v)
    assert EmulRegex(r"\<x").full_match(b"x") == []          # start of the value counts as "no word byte before"
    assert EmulRegex(r"x\>").full_match(b"x") == []          # end of the value counts as "no word byte next"
    assert EmulRegex(r"\>x").full_match(b"x") is None and EmulRegex(r"x\<").full_match(b"x") is None
    assert orc.Regex(r"[\<a]+").full_match(b"<a<") is not None    # inside a set the escape stays a literal


def test_single_byte_look_ahead_agrees_with_pcre2_and_python():
    """(?=x) / (?!x) with a one-byte body (literal, escape, class, '.'): an assertion on the next byte, compiled into
    every automaton -- full match with captures through the two-pass / forward tables and the single-pass tagged DFA,
    and the anchored prefix probe of the multiline patterns (the case the reference's users write: a start pattern like
    (?!\\s) for "the line does not begin with a blank")."""
    rng = random.Random(4711)
    pats = [r"(?!\s)(\S+) (\d+)(.*)", r"(\w+)(?=,)(.*)", r"a(?![bc])(.)(.*)", r"(?=\[)\[(\d+)\] (.*)", r"(\d+)(?!\d)(.*)",
            r"(.*?)(?=x)x(.*)", r"([a-c]*)(?!.)", r"(a|b(?=c))(.*)", r"(?i)(\w+) (?=Q)(.)(.*)", r"(\S*)(?![^ ])( ?)(.*)",
            r"(?=[a-c])(?!b)(\w+)(.*)", r"(x(?!y))*(.*)"]
    alpha = "abcxyq 1,[]Q"
    for p in pats:
        e, o, py = EmulRegex(p), orc.Regex(p), re.compile(p.encode(), re.S | re.M)
        assert e.supported, (p, e.error)
        for _ in range(500):
            v = "".join(rng.choice(alpha) for _ in range(rng.randint(0, 10))).encode()
            got, want = e.full_match(v), o.full_match(v)
            m = py.fullmatch(v)
            assert (got is None) == (want is None) == (m is None), (p, v, got, want)
            if got is not None:
                assert got == want, (p, v)
            if e.tdfa_info["states"] > 1:
                for mis in (0, 1):
                    assert e.full_match_tdfa(v, mis) == want, (p, v, mis)
            assert e.prefix_match(v) == o.prefix_match(v), (p, v)
    # the multiline use: start pattern "not a blank first"
    start = EmulRegex(r"(?!\s).*")
    assert start.supported
    assert start.prefix_match(b"Exception in thread") and not start.prefix_match(b"    at com.x.Y") \
        and not start.prefix_match(b"\tat z")
    assert start.prefix_match(b"") == orc.Regex(r"(?!\s).*").prefix_match(b"")


@pytest.mark.parametrize("pattern", [r"(", r"a)", r"[a", r"*a", r"a{2,1}", "a\\"])
def test_invalid_patterns(pattern):
    r = EmulRegex(pattern)
    assert not r.valid and not r.supported


# ----------------------------------------------------------------------------- random patterns
ATOMS = ["a", "b", "c", "x", " ", "-", "1", "2", r"\d", r"\w", r"\s", r"\S", r"\D", r"\W", ".", "[ab]", "[^a]", "[a-c1]",
         r"[\d\.]", r"[^\s\]]", r"\[", r"\]", r"\.", '"', "[^\"]", r"\t", r"[[:alpha:]]", r"\x41", ":", "/"]
QUANTS = ["", "", "", "*", "+", "?", "*?", "+?", "??", "{2}", "{1,3}", "{0,2}?", "{2,}"]


def gen(rng, depth=0):
    n = rng.randint(1, 4)
    parts = []
    for _ in range(n):
        r = rng.random()
        if depth < 3 and r < 0.25:
            inner = gen(rng, depth + 1)
            if rng.random() < 0.35:
                inner = inner + "|" + gen(rng, depth + 1)
                if rng.random() < 0.2:
                    inner += "|"
            atom = ("(" if rng.random() < 0.75 else "(?:") + inner + ")"
        else:
            atom = rng.choice(ATOMS)
        parts.append(atom + rng.choice(QUANTS))
    if depth == 0 and rng.random() < 0.15:
        parts.insert(0, "^")
    if depth == 0 and rng.random() < 0.15:
        parts.append("$")
    if rng.random() < 0.05:
        parts.insert(rng.randint(0, len(parts)), r"\b")
    if rng.random() < 0.04:
        parts.insert(rng.randint(0, len(parts)), rng.choice(["^", "$"]))
    return "".join(parts)


ALPHA = "abcx  --12\t\"[].:/A_\n"


def rand_input(rng, alpha):
    return "".join(rng.choice(alpha) for _ in range(rng.randint(0, 14))).encode()


def _bol_at_end_corner(pat, s):
    return "^" in pat[1:] and s[-1:] in (b"\n", b"\r")


def py_caps(m, n):
    out = []
    for g in range(1, m.re.groups + 1):
        s, e = m.span(g)
        out.append((n, 0) if s < 0 else (s, e - s))
    return out


@pytest.mark.parametrize("seed", range(8))
def test_random_patterns_agree_with_pcre2_and_python(seed):
    rng = random.Random(20260922 + seed)
    checked = n_sup = n_two = n_tdfa = 0
    for _ in range(400):
        pat = gen(rng)
        try:
            o = orc.Regex(pat)
        except ValueError:
            continue
        e = EmulRegex(pat)
        assert e.valid, (pat, e.error)
        if not e.supported:
            assert "empty string" in e.error or "too large" in e.error, (pat, e.error)
            continue
        n_sup += 1
        n_two += e.mode
        assert e.ngroups == o.ngroups, pat
        try:
            pr = re.compile(pat.encode(), re.S | re.M)
        except re.error:
            pr = None
        if pr is not None and re.search(rb"\\[vhVH]|\[\[:", pat.encode()):
            pr = None  # python spells these classes differently
        for _ in range(25):
            s = rand_input(rng, ALPHA)
            want = o.full_match(s)
            got = e.full_match(s)
            td = e.full_match_tdfa(s, rng.randint(0, 15))
            assert td == "n/a" or td == got, ("tdfa", pat, s, td, got)
            n_tdfa += td != "n/a"
            if _bol_at_end_corner(pat, s):
                # boost (and Python) let a mid-pattern '^' match at END of input after a trailing newline;
                # PCRE2 does not.  The engine follows boost; only Python can arbitrate this corner.
                if pr is not None:
                    m = pr.fullmatch(s)
                    assert (py_caps(m, len(s)) if m else None) == got, (pat, s)
                    assert (pr.match(s) is not None) == e.prefix_match(s), (pat, s)
                continue
            assert got == want, (pat, s, got, want)
            assert e.prefix_match(s) == o.prefix_match(s), (pat, s)
            f2 = e.full_match_fast2(s, rng.randint(0, 15))
            assert f2 == "n/a" or f2 == want, ("fast2", pat, s, f2, want)
            if pr is not None:
                m = pr.fullmatch(s)
                assert (m is None) == (want is None), (pat, s)
                if m is not None:
                    assert py_caps(m, len(s)) == want, (pat, s)
                assert (pr.match(s) is not None) == e.prefix_match(s), (pat, s)
            checked += 1
    assert checked > 2000 and n_sup > 100
    assert n_tdfa > checked // 2  # the single-pass layout must cover most patterns


def test_random_crlf_line_anchors_agree_with_pcre2_anycrlf():
    """boost treats \\n, \\r (and \\r\\n as a unit) as line separators for ^ and $;
    PCRE2's ANYCRLF convention is the same rule set minus \\f, so inputs here avoid \\f."""
    rng = random.Random(77)
    alpha = "ab \r\n\r\n"
    n = 0
    for _ in range(600):
        pat = gen(rng)
        if "^" not in pat and "$" not in pat:
            continue
        try:
            o = orc.Regex("(*ANYCRLF)" + pat)
        except ValueError:
            continue
        e = EmulRegex(pat)
        if not e.supported:
            continue
        for _ in range(30):
            s = rand_input(rng, alpha)
            if _bol_at_end_corner(pat, s) or b"\r\n" in s:
                continue  # boost never anchors BETWEEN \r and \n; PCRE2 does -- see test_boost_crlf_unit
            assert e.full_match(s) == o.full_match(s), (pat, s)
            td = e.full_match_tdfa(s, rng.randint(0, 127))
            assert td == "n/a" or td == e.full_match(s), ("tdfa", pat, s)
            assert e.prefix_match(s) == o.prefix_match(s), (pat, s)
            n += 1
    assert n > 500


def test_boost_crlf_unit():
    """perl_matcher::match_start_line / match_end_line (Boost.Regex 1.68, perl_matcher_common.hpp): a \\r\\n pair is
    ONE separator: '$' matches before the \\r, '^' after the \\n, neither in between.  \\f is a separator too."""
    e = EmulRegex(r"a$.^b")
    assert e.full_match(b"a\nb") == []
    assert e.full_match(b"a\rb") == []
    assert e.full_match(b"a\fb") == []
    assert e.full_match(b"a b") is None
    e = EmulRegex(r"a$..^b")
    assert e.full_match(b"a\r\nb") == []
    e = EmulRegex(r"a.$.^b")
    assert e.full_match(b"a\r\nb") is None      # '$' between \r and \n
    assert e.full_match(b"a\n\nb") == []
    e = EmulRegex(r"a$.^.b")
    assert e.full_match(b"a\r\nb") is None      # '^' between \r and \n
    assert e.full_match(b"a\n\rb") == []
    e = EmulRegex(r"a\n^")
    assert e.full_match(b"a\n") == []            # '^' at end of input after a trailing separator (boost, Python)


def test_tdfa_layout_on_benchmark_patterns_and_corners():
    """The single-pass tagged DFA (the headline kernel's tables) against the oracle on realistic lines, at both
    pair alignments, plus the corners its register scheme must get right: empty groups (two boundaries in one
    step -> slow path), optional groups left unset, captures inside repeats (last iteration wins) and
    alternations whose losing branch wrote a register first."""
    from loongcollector_b200 import synth
    rng = random.Random(5)
    for pat in (synth.NGINX_PATTERN, synth.APACHE_PATTERN, synth.CSV_URL_PATTERN):
        e, o = EmulRegex(pat), orc.Regex(pat)
        info = e.tdfa_info
        assert info["states"] > 1 and info["regs"] >= 2 * e.ngroups, (pat, info)
        buf, off, ln = synth.nginx_lines(300, seed=rng.randint(0, 1 << 30), line_bytes=None)
        lines = [bytes(buf[a:a + b]) for a, b in zip(off, ln)]
        lines += [l[:rng.randint(0, len(l))] for l in lines[:60]] + [l.replace(b'"-"', b'""') for l in lines[:60]]
        for l in lines:
            want = o.full_match(l)
            for mis in (0, 1, 14, 15, 77):
                assert e.full_match_tdfa(l, mis) == want, (pat, l, mis)
    corners = [
        (r"(a*)(b*)(c*)", [b"", b"a", b"b", b"c", b"abc", b"aacc", b"bb"]),
        (r"(?:(a)|(b)|(c))*", [b"", b"a", b"ab", b"abc", b"cba", b"aab"]),
        (r"(a|ab)(c|bcd)(d*)", [b"abcd", b"acd", b"abcdd", b"ac"]),
        (r"(\d+)-(\d+)?(x|yy)*(.*?)(\s.*|)", [b"1-2xyy z", b"12-", b"1-xx", b"3-4yyx\tq"]),
        (r'"([^"]*)" "([^"]*)"', [b'"" ""', b'"a" ""', b'"" "b"', b'"ab" "cd"']),
        (r"(\S+\]) (x*)", [b"a]b] xx", b"]]] ", b"a] x"]),
        (r"^(\w+)\b.(\w+)$", [b"ab cd", b"ab,cd", b"a\nb"]),
    ]
    for pat, inputs in corners:
        e, o = EmulRegex(pat), orc.Regex(pat)
        assert e.tdfa_info["states"] > 1, pat
        for s in inputs:
            for mis in (0, 1):
                assert e.full_match_tdfa(s, mis) == o.full_match(s), (pat, s, mis)


def _tdfa_blob(e):
    import ctypes as C
    import struct

    import numpy as np
    from tests.emul.emul import lib
    L = lib()
    L.emul_tdfa_blob.restype = C.c_uint32
    L.emul_tdfa_blob.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    n = L.emul_tdfa_blob(e._h, None, 0)
    if not n:
        return None, None
    buf = np.zeros(n, np.uint8)
    L.emul_tdfa_blob(e._h, buf.ctypes.data_as(C.c_void_p), n)
    b = buf.tobytes()
    names = ("magic total ngroups nstates ncls nregs start row_bytes off_cls off_t2 off_t1 off_eof off_ops has_slow "
             "max_threads sink").split()
    return dict(zip(names, struct.unpack("<16I", b[:64]))), b


def test_tdfa_table_invariants_the_kernel_relies_on():
    """Structural facts the staged kernel assumes about every single-pass blob (lc_tables.h: LcTdfaHeader): rows fit
    16-bit offsets even after rebasing, the row pitch is an odd word count, the dead row and the sink row absorb, every
    entry points at a row start, register fields name existing registers, slow entries lead to the sink, and the
    op lists stay inside the pool."""
    import numpy as np
    rng = random.Random(31)
    seen = slow = 0
    pats = [gen(rng) for _ in range(500)] + [r"(\w+) (\d+)", r'"([^"]*)" "([^"]*)"', r"(?:(a)|(b)|(c))*x"]
    from loongcollector_b200 import synth
    pats += [synth.NGINX_PATTERN, synth.APACHE_PATTERN, synth.JAVA_START_PATTERN, synth.CSV_URL_PATTERN]
    for pat in pats:
        e = EmulRegex(pat)
        if not e.supported:
            continue
        h, b = _tdfa_blob(e)
        if h is None:
            continue
        seen += 1
        ns, ncl, rb = h["nstates"], h["ncls"], h["row_bytes"]
        assert h["magic"] == 0x4C435444 and h["total"] == len(b) and h["sink"] == ns and h["start"] == 1
        assert rb == ((ncl * ncl) | 1) * 4 and (rb // 4) % 2 == 1
        assert (ns + 2) * rb + 2048 <= 65535 and h["nregs"] <= 62 and h["nregs"] >= 2 * h["ngroups"]
        t2 = np.frombuffer(b[h["off_t2"]:h["off_t2"] + (ns + 1) * rb], np.uint32).reshape(ns + 1, rb // 4)
        used = t2[:, :ncl * ncl]
        nxt = used & 0xFFFF
        assert np.all(nxt % rb == 0) and np.all(nxt // rb <= ns)
        assert np.all(t2[0] == 0)                                       # dead row absorbs, no actions
        assert np.all(used[ns] == (ns * rb | 0x00800000))               # sink row absorbs
        is_slow = (used & 0x00800000) != 0
        assert np.all((nxt[is_slow] // rb) == ns)                       # slow entries lead to the sink ...
        assert np.all((used[is_slow] & 0x7F7F0000) == 0)                # ... and carry no register fields
        assert bool(is_slow[:ns].any()) == bool(h["has_slow"])
        slow += int(h["has_slow"])
        for shift in (16, 24):
            f = (used[:ns] >> shift) & 0x7F
            f = f[(f != 0) & ~is_slow[:ns]]
            assert np.all(f % 2 == 0) and np.all((f - 2) // 2 < h["nregs"])
        t1 = np.frombuffer(b[h["off_t1"]:h["off_t1"] + ns * ncl * 4], np.uint32)
        ops = np.frombuffer(b[h["off_ops"]:], np.uint16)
        eof = np.frombuffer(b[h["off_eof"]:h["off_eof"] + ns * 4], np.uint32)
        assert np.all((t1 & 0xFFFF) < ns)
        for lst in list(t1 >> 16) + [x for x in eof if x != 0xFFFFFFFF]:
            lst = int(lst)
            assert lst < ops.size and lst + 1 + int(ops[lst]) <= ops.size
            for op in ops[lst + 1:lst + 1 + int(ops[lst])]:
                dst, src = int(op) >> 8, int(op) & 0xFF
                assert dst < h["nregs"] and (src in (0xFF, 0xFE) or src < h["nregs"])
    assert seen > 100 and slow > 5


def test_tdfa_run_skipping_chunks_is_exact():
    """States that loop on every byte but (at most) two exit bytes jump over whole 16-byte chunks without those bytes
    (lc_tables.h: skip).  Fields of every length around the chunk size, exit bytes at every slot of a chunk, at every
    alignment of the line -- against the oracle."""
    pats = [r'"([^"]*)" (.*)', r'(\w+) "([^\\"]*)" (\d+)', r"\[([^\]]+)\] (.*)", r"(.*)", r'x(.*)y"([^"]*)"']
    for pat in pats:
        e, o = EmulRegex(pat), orc.Regex(pat)
        assert e.supported and e.tdfa_info["states"] > 1, pat
        cases = []
        for n in list(range(0, 40)) + [47, 48, 49, 63, 64, 65, 100, 255]:
            fill = ("ab/c?d=" * 40)[:n]
            cases += [('"%s" tail %s' % (fill, fill)).encode(), ('GET "%s" 200' % fill).encode(),
                      ("[%s] rest%s" % (fill, fill)).encode(), ("x%sy\"%s\"" % (fill, fill)).encode(),
                      ('"%s\\" 1' % fill).encode(), ('"%s' % fill).encode()]
        for v in cases:
            want = o.full_match(v)
            for mis in range(16):
                assert e.full_match_tdfa(v, mis) == want, (pat, v, mis)
