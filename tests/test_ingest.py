"""Raw ingest (ga_parse_reads / ga_ingest.parse) against the text parser it replaces: same reads, kind,
distance and base count as IOHandler.read_input's text path for every rule of SURVEY App. A-17
(assemble.py:40-71 of the reference).  Host-only: no GPU needed."""
import io

import pytest

import assemble
import ga_ingest


def _text_path(text: str):
    return assemble.IOHandler.read_input(io.StringIO(text, newline=None))     # universal newlines, as sys.stdin


CASES = [
    "3\nACGT\nGGCA\nTTTT\n",
    "3\nACGT\nGGCA\nTTTT",                          # no trailing newline
    "2\r\nACGT\r\nGGCA\r\n",                        # CRLF
    "2\rACGT\rGGCA\r",                              # lone CR (universal newlines)
    " 2 \n  ACGT \t\n\tGG CA\x0b\n",                # white space around lines, kept inside
    "4\nACGT\nGG\n",                                # missing lines are empty reads
    "0\nACGT\nGGCA\n",                              # n = 0 still consumes one read
    "1\nACGT\nGGCA\nTTTT\n",                        # trailing lines ignored
    "2\nACGT|TTGA|125\nGGCA|CCAT|125\n",
    "2\n ACGT|TTGA|125 \r\nGGCA|CCAT| 7 \r\n",
    "1\nAC_GT;|x y|3\n",
    "3\n\n\n\n",
    "2\nnnabe_Lee;|_Lee;By_th|5\nBy_the_nam|e_name_of_|5\n",
    "+2\nAC\nGT\n",
]


@pytest.mark.parametrize("text", CASES)
def test_raw_parse_matches_text_parse(text):
    want = _text_path(text)
    got = ga_ingest.parse(text.encode("ascii"))
    assert got is not None
    reads, paired, distance, bases = got
    assert (list(reads), paired, distance, bases) == (list(want[0]), want[1], want[2], want[3])
    assert len(reads) == len(want[0])
    if len(reads):
        assert reads[0] == want[0][0] and reads[-1] == want[0][-1]
    # the CLI entry takes the raw route when it is handed bytes
    via_cli = assemble.IOHandler.read_input(io.BytesIO(text.encode("ascii")))
    assert (list(via_cli[0]), via_cli[1], via_cli[2], via_cli[3]) == (list(want[0]), want[1], want[2], want[3])


@pytest.mark.parametrize("text", ["2\nACGT|TT|1\nGGCA\n", "2\nAC|GT|1|9\nAA|CC|1\n", "3\nAC|GT|1\nAA|CC|1\n"])
def test_raw_parse_rejects_malformed_pairs_like_the_reference(text):
    with pytest.raises(ValueError):
        _text_path(text)
    with pytest.raises(ValueError):
        ga_ingest.parse(text.encode("ascii"))


def test_raw_parse_defers_to_text_for_what_it_does_not_cover():
    assert ga_ingest.parse("2\nACéT\nGG\n".encode("utf-8")) is None     # not plain ASCII
    assert ga_ingest.parse(b"two\nAC\nGG\n") is None                          # header is not an integer
    assert ga_ingest.parse(b"") is None
    with pytest.raises(ValueError):
        assemble.IOHandler.read_input(io.BytesIO(b"two\nAC\nGG\n"))           # Python's own error from the text path
    reads, paired, _, _ = assemble.IOHandler.read_input(io.BytesIO("1\nACéT\n".encode("utf-8")))
    assert list(reads) == ["ACéT"] and not paired


def test_raw_reads_sequence_surface():
    reads, paired, distance, bases = ga_ingest.parse(b"3\nAAAA|CCCC|9\nGG|TT|9\nA|C|9\n")
    assert paired and distance == 9 and bases == 14
    assert len(reads) == 3 and reads[1] == ("GG", "TT") and reads[-1] == ("A", "C")
    assert reads[0:2] == [("AAAA", "CCCC"), ("GG", "TT")]
    assert [r for r in reads] == [("AAAA", "CCCC"), ("GG", "TT"), ("A", "C")]
    with pytest.raises(IndexError):
        reads[3]


THREADED = [c for c in CASES if "\r" not in c] + [
    "7\nACGT\nGG\n\nTTTTTTTT\nA\nCC\nGGG\nextra\nlines\n",
    "5\nAC|GT|3\nAAAA|CCCC|3\nG|T|3\nGGGG|TTTT|3\nA|C|41\nXX|YY|9\n",
    "9\nACGT\nGG\n",                                # more reads announced than lines present
    "1\n" + "ACGT" * 300 + "\n",
]


@pytest.mark.parametrize("threads", [2, 3, 8])
@pytest.mark.parametrize("text", THREADED)
def test_parallel_parse_equals_serial_parse(text, threads, monkeypatch):
    """ga_parse_reads cuts big inputs into one byte range per thread; forced here on small inputs (one-byte grain):
    same reads, kind, distance and base count as the serial scan and as the text parser."""
    monkeypatch.setenv("GA_PARSE_THREADS", "1")
    serial = ga_ingest.parse(text.encode("ascii"))
    monkeypatch.setenv("GA_PARSE_THREADS", str(threads))
    monkeypatch.setenv("GA_PARSE_GRAIN", "1")
    got = ga_ingest.parse(text.encode("ascii"))
    want = _text_path(text)
    for reads, paired, distance, bases in (serial, got):
        assert (list(reads), paired, distance, bases) == (list(want[0]), want[1], want[2], want[3])


@pytest.mark.parametrize("text", ["2\nACGT|TT|1\nGGCA\n", "2\nAC|GT|1|9\nAA|CC|1\n", "3\nAC|GT|1\nAA|CC|1\n",
                                  "2\nAA|CC|1\nAC|GT|x\n"])
def test_parallel_parse_rejects_what_the_serial_parse_rejects(text, monkeypatch):
    monkeypatch.setenv("GA_PARSE_THREADS", "4")
    monkeypatch.setenv("GA_PARSE_GRAIN", "1")
    with pytest.raises(ValueError):
        _text_path(text)
    if text.endswith("|x\n"):       # a distance only Python can judge: the raw parser hands the input to the text path
        assert ga_ingest.parse(text.encode("ascii")) is None
    else:
        with pytest.raises(ValueError):
            ga_ingest.parse(text.encode("ascii"))
    with pytest.raises(ValueError):
        assemble.IOHandler.read_input(io.BytesIO(text.encode("ascii")))


def test_parallel_parse_large_input(monkeypatch):
    """A few MB through the default thread count: equal to the serial scan byte for byte."""
    import numpy as np
    rng = np.random.default_rng(5)
    n = 40000
    lens = rng.integers(0, 180, n)
    body = b"".join(bytes(rng.integers(65, 91, int(m), dtype=np.uint8)) + b"\n" for m in lens)
    text = (b"%d\n" % n) + body
    monkeypatch.setenv("GA_PARSE_THREADS", "1")
    a = ga_ingest.parse(text)
    monkeypatch.setenv("GA_PARSE_THREADS", "6")
    monkeypatch.setenv("GA_PARSE_GRAIN", "65536")
    b = ga_ingest.parse(text)
    assert a[1:] == b[1:] and np.array_equal(a[0].lens, b[0].lens) and np.array_equal(a[0].symbols, b[0].symbols)
    assert a[0].lens.tolist() == lens.tolist()


def _reference_read_input(raw: bytes, encoding: str = "ascii"):
    """The UNMODIFIED reference's IOHandler.read_input (oracle/_ref, copied from /root/reference by oracle/make_ref.py)
    over `raw` as its stdin: ("ok", reads, paired, distance, bases) or ("err", exception name)."""
    import importlib.util
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    ref_dir = os.path.join(os.path.dirname(here), "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "assemble.py")):
        pytest.skip("oracle/_ref is only present where /root/reference is")
    if "_ref_assemble" not in sys.modules:
        sys.path.insert(0, ref_dir)
        try:
            spec = importlib.util.spec_from_file_location("_ref_assemble", os.path.join(ref_dir, "assemble.py"))
            module = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(module)
            sys.modules["_ref_assemble"] = module
        finally:
            sys.path.remove(ref_dir)
            for name in ("debruijn_graph", "debug_graph", "debruijn_node", "countminsketch"):
                mod = sys.modules.get(name)          # the reference's siblings must not shadow the product's modules
                if mod is not None and getattr(mod, "__file__", "").startswith(ref_dir):
                    del sys.modules[name]
    ref = sys.modules["_ref_assemble"]
    old = sys.stdin
    sys.stdin = io.TextIOWrapper(io.BytesIO(raw), encoding=encoding)
    try:
        return ("ok",) + tuple(ref.IOHandler.read_input())
    except Exception as exc:        # noqa: BLE001 -- the exception TYPE is what is compared
        return ("err", type(exc).__name__)
    finally:
        sys.stdin = old


@pytest.mark.parametrize("threads", [1, 3])
def test_fuzz_against_the_unmodified_reference(threads, monkeypatch):
    """Adversarial stdin texts (odd headers, white space of every kind Python strips, CR / CRLF / LF, empty and
    missing lines, malformed pair lines, distance fields only Python's int() understands) through the product's
    IOHandler.read_input and through the reference's own: same reads, kind, distance, base count -- or the same
    exception type."""
    import random
    monkeypatch.setenv("GA_PARSE_THREADS", str(threads))
    monkeypatch.setenv("GA_PARSE_GRAIN", "1")
    rng = random.Random(20261019 + threads)
    alphabet = "ACGTNacgt_;x"
    spaces = [" ", "\t", "\x0b", "\x0c", "\x1c", "\x1d", "\x1e", "\x1f"]

    def read():
        s = "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 8)))
        if rng.random() < 0.1:
            at = rng.randrange(len(s) + 1)
            s = s[:at] + rng.choice(spaces) + s[at:]
        return s

    for _ in range(2500):
        n = rng.randrange(0, 6)
        paired = rng.random() < 0.5
        lines = []
        for _line in range(rng.randrange(0, 7)):
            if paired and rng.random() < 0.93:
                parts = [read(), read(), rng.choice(["125", " 7", "3 ", "-2", "+4", "x", "", "1_0", "0x10", "1e3"])]
                if rng.random() < 0.05:
                    parts.append("9")
                if rng.random() < 0.05:
                    parts = parts[:2]
                line = "|".join(parts)
            else:
                line = read()
            if rng.random() < 0.15:
                line = rng.choice(spaces) + line
            if rng.random() < 0.15:
                line = line + rng.choice(spaces)
            lines.append(line)
        head = rng.choice([str(n), " %d " % n, "+%d" % n, "-%d" % n, "0%d" % n, "%d.0" % n, "", "1_0", str(n) + "\t"])
        nl = rng.choice(["\n", "\n", "\r\n", "\r"])
        raw = (head + nl + nl.join(lines) + (nl if rng.random() < 0.7 else "")).encode("ascii")
        want = _reference_read_input(raw)
        try:
            got = assemble.IOHandler.read_input(io.BytesIO(raw))
            got = ("ok", list(got[0]), got[1], got[2], got[3])
        except Exception as exc:        # noqa: BLE001
            got = ("err", type(exc).__name__)
        assert got == want, raw


def test_fuzz_unicode_against_the_unmodified_reference():
    """The same with UTF-8 input -- letters outside ASCII, white space only str.strip() knows (NBSP, U+2003, U+0085,
    U+3000), digits only int() knows (Arabic-Indic, full-width), a stray 0xFF byte: the raw parser steps aside and the
    text path must apply Python's rules exactly as the reference does."""
    import random
    rng = random.Random(20261020)
    alphabet = "ACGT\u00e9\u0663x"
    spaces = [" ", "\t", "\u00a0", "\u2003", "\u0085", "\u2009", "\x1c", "\u3000"]

    def read():
        s = "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 7)))
        if rng.random() < 0.15:
            at = rng.randrange(len(s) + 1)
            s = s[:at] + rng.choice(spaces) + s[at:]
        return s

    for _ in range(2500):
        n = rng.randrange(0, 5)
        paired = rng.random() < 0.5
        lines = []
        for _line in range(rng.randrange(0, 6)):
            if paired and rng.random() < 0.93:
                line = "|".join([read(), read(), rng.choice(["125", " 7", "\u0663", "\u00a05", "x", "", "\uff11\uff12"])])
            else:
                line = read()
            if rng.random() < 0.2:
                line = rng.choice(spaces) + line
            if rng.random() < 0.2:
                line = line + rng.choice(spaces)
            lines.append(line)
        head = rng.choice([str(n), " %d " % n, "\u0663", "\u00a0%d" % n, "%d\u2003" % n, ""])
        nl = rng.choice(["\n", "\n", "\r\n", "\r"])
        raw = (head + nl + nl.join(lines) + (nl if rng.random() < 0.7 else "")).encode("utf-8")
        if rng.random() < 0.03:
            at = rng.randrange(len(raw) + 1)
            raw = raw[:at] + b"\xff" + raw[at:]
        want = _reference_read_input(raw, "utf-8")
        try:
            got = assemble.IOHandler.read_input(io.BytesIO(raw))
            got = ("ok", list(got[0]), got[1], got[2], got[3])
        except Exception as exc:        # noqa: BLE001
            got = ("err", type(exc).__name__)
        assert got == want, raw
