"""SURVEY 8(f) rank 3 -- the YAML scene format (reference scene.c:573-2175, scenes/yaml.c).

The reference delegates all YAML syntax to libyaml's event API; include/yaml_lite/yaml.h +
ndt_b200/csrc/yaml_lite.c provide that API without libyaml, and the reference's scene.c compiles
against it UNMODIFIED (oracle/Makefile, -DWITH_YAML).  So parity has two halves:

 * the event streams and the emitted text are pinned against libyaml 0.2.5 ITSELF (PyYAML's
   CParser / CEmitter embed it) -- fixtures, hand-written cases and seeded random documents;
 * scenes loaded from YAML render identically: the `yaml` cases of tests/scenes.py go through the
   same golden / oracle / emulation / GPU tiers as every other scene (test_oracle.py,
   test_gpu_parity.py); here: the reference's own emitter and parser running over yaml_lite.
"""
import os
import random
import re
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from yaml_lite_util import LIB, ROOT, lite_events, lite_emit, libyaml_events, libyaml_emit  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refharness  # noqa: E402

SCENES = os.path.join(ROOT, "tests", "scenes")

HAND = [
    b"a: 1\nb:\n  - x\n  - y: 2\n    z: [1, 2, {k: v}]\nc: 'it''s'\nd: \"q\\n\\x41\\u00e9\"\n# comment\ne:   # trailing\n  f: g #c\n",
    b"--- \nscene: s\n...\n---\nscene: t\n",
    b"- - a\n  - b\n- c\n-\n  d: e\n- \n",
    b"k: [1,\n  2,\n 3]\nm: {a: 1,\n b: 2}\n",
    b"",
    b"# only a comment\n",
    b"--- text\n--- [a]\n--- {a: b}\n",
    b"a:\n- b\n- c\nd: e\n",
    b"x: {a, b: , c: d}\ny: [a: b, c]\n",
    b"key with spaces: value with spaces   \nurl: http://x.y:80/z#frag\nneg: -1.5e-3\ndash: -x\n",
    b"'quoted key': \"dq\"\n\"k2\": 'folded\n  over two lines'\nk3: \"escaped \\\n   break\"\n",
    b"a: \"two\n\n  breaks\"\n",
    b"\xef\xbb\xbfbom: 1\n",
    b"a: 1\r\nb: [1,\r\n 2]\r\n",
    b"---\na: 1\n...\n...\n---\nb: 2\n",
    b"a: {}\nb: []\nc: [[], {}]\n",
    b"sizes:\n- 5.0\n-   6\nflags: [ 0 ]\n",
    b"- a\n b\n",
    b"? a\n: b\n? [c, d]\n: - e\n  - f\n? g\n\"\": h\nx: {? i : j, ? k, ? : l}\n",
    b"name: a long name\n  that continues\n\n  after a blank line   # comment\nnext: [one\n  two, three]\n--- top level\nplain\n",
]
BAD = [
    b"a: b: c\n",
    b"a: &x 1\n",
    b"a: *x\n",
    b"a: !!str 1\n",
    b"a: |\n  text\n",
    b"%YAML 1.1\n---\na: 1\n",
    b"a: [1, 2\nb: 3\n",
    b"a: 'open\n",
    b"a: 1\n  b: 2\n",
    b"a:\n\t- 1\n",
    b"...\na: 1\n",
    b"a: 1\n...\nb: 2\n",
]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "yaml_lite", "yaml.h")).read()
    names = set(re.findall(r"#define\s+yaml_\w+\s+(ylite_\w+)", hdr)) | set(re.findall(r"\b(ylite_\w+)\s*\(", hdr))
    assert len(names) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True, check=True).stdout
    have = {l.split()[-1] for l in out.splitlines() if l.strip()}
    assert not (names - have), f"declared but not exported: {sorted(names - have)}"


@pytest.mark.parametrize("name", ["config5_mixed10d.yaml", "handwritten4d.yaml"])
def test_parser_events_equal_libyaml_on_scene_files(name):
    text = open(os.path.join(SCENES, name), "rb").read()
    rc, mine = lite_events(text)
    err, theirs = libyaml_events(text)
    assert rc == 0 and err == 0
    assert mine == theirs
    assert mine.count("+DOC") == (3 if name.startswith("hand") else 1)


@pytest.mark.parametrize("name", ["config5_mixed10d.yaml", "handwritten4d.yaml"])
def test_parser_reads_the_layouts_of_another_emitter(name):
    """the same scenes re-written by PyYAML's own (pure Python) emitter in five layouts -- all block, all flow,
    4-space indent at width 60, explicit document markers at width 1000: same events as libyaml reads"""
    import yaml
    docs = list(yaml.safe_load_all(open(os.path.join(SCENES, name))))
    for kw in (dict(default_flow_style=None), dict(default_flow_style=False), dict(default_flow_style=True),
               dict(default_flow_style=None, indent=4, width=60),
               dict(default_flow_style=None, explicit_start=True, explicit_end=True, width=1000)):
        text = yaml.safe_dump_all(docs, **kw).encode()
        rc, mine = lite_events(text)
        err, theirs = libyaml_events(text)
        assert rc == 0 and err == 0 and mine == theirs, kw


@pytest.mark.parametrize("i", range(len(HAND)))
def test_parser_events_equal_libyaml_handwritten(i):
    rc, mine = lite_events(HAND[i])
    err, theirs = libyaml_events(HAND[i])
    assert err == 0, "libyaml rejects the test input"
    assert rc == 0, mine
    assert mine == theirs


@pytest.mark.parametrize("i", range(len(BAD)))
def test_parser_refuses_what_it_does_not_support(i):
    """Out-of-subset or malformed input is an error (never a guess), and every event handed out before
    the error is one libyaml also produced."""
    rc, mine = lite_events(BAD[i])
    assert rc in (3, 4), mine                       # YAML_SCANNER_ERROR / YAML_PARSER_ERROR
    lines = mine.splitlines()
    assert lines[-1].startswith("!ERR")
    _, theirs = libyaml_events(BAD[i])
    t = theirs.splitlines()
    # the two report at slightly different events (libyaml folds a multi-line plain scalar before it
    # notices, yaml_lite refuses at the line break), so compare the common prefix short of the last event
    k = min(len(lines) - 2, len(t))
    assert lines[:k] == t[:k]


def test_flow_collection_as_a_simple_key_is_refused():
    """valid YAML that libyaml reads and yaml_lite does not: an error, not a guess"""
    for text in (b"[a, b]: c\n", b"{a: b}: c\n", b"[]: c\n"):
        rc, mine = lite_events(text)
        assert rc == 4 and "flow collections as mapping keys" in mine.splitlines()[-1]
        assert libyaml_events(text)[0] == 0


ALPHA = "ab1 -:#,[]{}'\"?!&*|>%@`.~\\\t\n=<x y"


def _rstr(rng, empty_ok=True):
    n = rng.choice([0 if empty_ok else 1, 1, 1, 2, 3, 5, 8, 20, 60, 150])
    return "".join(rng.choice(ALPHA) for _ in range(n))


def _esc(v):
    return v.replace("\\", "\\\\").replace("\n", "\\n").replace("\t", "\\t")


def _node(rng, L, depth, quoted_implicit_only, keys_simple):
    r = rng.random()

    def val(s):
        p = rng.choice("01")
        q = "1" if (p == "0" or quoted_implicit_only) else rng.choice("01")
        return "=VAL %s%sa %s" % (p, q, _esc(s))
    if depth > 3 or r < 0.5:
        L.append(val(_rstr(rng)))
    elif r < 0.75:
        L.append("+SEQ" + rng.choice(["", " []"]))
        for _ in range(rng.randint(0, 4)):
            _node(rng, L, depth + 1, quoted_implicit_only, keys_simple)
        L.append("-SEQ")
    else:
        L.append("+MAP" + rng.choice(["", " {}"]))
        for _ in range(rng.randint(0, 4)):
            if keys_simple:
                s = _rstr(rng, empty_ok=False).replace("\n", " ")[:100]
                L.append(val(s))
            elif rng.random() < 0.85:
                L.append(val(_rstr(rng)))
            else:
                _node(rng, L, depth + 1, quoted_implicit_only, keys_simple)
            _node(rng, L, depth + 1, quoted_implicit_only, keys_simple)
        L.append("-MAP")


def _stream(rng, quoted_implicit_only=False, keys_simple=False):
    L = ["+STR"]
    for _ in range(rng.choice([1, 1, 2, 3])):
        L.append("+DOC" + rng.choice(["", " ---"]))
        _node(rng, L, 0, quoted_implicit_only, keys_simple)
        L.append("-DOC" + rng.choice(["", " ..."]))
    L.append("-STR")
    return "\n".join(L) + "\n"


def test_emitter_text_equals_libyaml_on_random_event_streams():
    """scalar style selection, quoting and escaping, block / flow layout, indentless sequences, empty
    collections, complex keys, the '!' tag of non-plain plain-implicit scalars, 80-column folding"""
    rng = random.Random(20261018)
    for t in range(1200):
        listing = _stream(rng)
        width = rng.choice([None, None, 30, 200])
        rc, mine = lite_emit(listing, width or 0)
        theirs = libyaml_emit(listing, width)
        assert rc == 0 and mine == theirs, f"stream {t}:\n{listing}\n--- libyaml\n{theirs!r}\n--- yaml_lite\n{mine!r}"


def test_emitter_folds_long_vectors_like_libyaml():
    """what scene_yaml_emit_vect (scene.c:713) produces for N-D vectors of %.16g doubles"""
    rng = random.Random(5)
    for n in (3, 4, 10, 16, 33):
        L = ["+STR", "+DOC ---", "+MAP"]
        for k in ("viewPoint", "pos"):
            L += ["=VAL 11a " + k, "+SEQ []"]
            L += ["=VAL 11a %.16g" % (rng.uniform(-100, 100) * rng.choice([1, 1e-9, 1e12])) for _ in range(n)]
            L += ["-SEQ"]
        L += ["=VAL 11a positions", "+SEQ"]
        for _ in range(3):
            L += ["+SEQ []"] + ["=VAL 11a %.16g" % rng.uniform(-10, 10) for _ in range(n)] + ["-SEQ"]
        L += ["-SEQ", "-MAP", "-DOC", "-STR"]
        listing = "\n".join(L) + "\n"
        rc, mine = lite_emit(listing)
        assert rc == 0 and mine == libyaml_emit(listing)
        if n >= 10:
            assert b",\n  " in mine                   # folded after column 80, continuation indented


def test_parser_events_equal_libyaml_on_random_documents():
    """libyaml writes random event streams (four widths: folded plain scalars, wrapped flow collections, complex
    "? " keys for empty / long / multi-line / collection keys); both parsers read them back.  The one construct
    yaml_lite refuses here is an empty flow collection used as a simple key ("[]: v"): refused, not misread."""
    rng = random.Random(77)
    refused = unreadable = 0
    for t in range(1000):
        listing = _stream(rng, quoted_implicit_only=True, keys_simple=(t % 2 == 0))
        text = libyaml_emit(listing, rng.choice([None, 30, 40, 1000]))
        err, theirs = libyaml_events(text)
        if err:                 # libyaml cannot read back everything it writes (e.g. a tab after '- ')
            unreadable += 1
            continue
        rc, mine = lite_events(text)
        if rc:
            refused += 1
            assert rc in (3, 4) and "flow collections as mapping keys" in mine.splitlines()[-1], mine.splitlines()[-1]
            continue
        assert mine == theirs, f"document {t}:\n{text.decode()}"
    assert refused < 30 and unreadable < 80, (refused, unreadable)


def test_roundtrip_through_both_directions():
    """emit -> parse -> emit is a fixed point (a size-independent property, here on a big stream)"""
    rng = random.Random(9)
    L = ["+STR", "+DOC ---", "+MAP", "=VAL 11a objects", "+SEQ"]
    for i in range(5000):
        L += ["+MAP", "=VAL 11a type", "=VAL 11a sphere", "=VAL 11a positions", "+SEQ", "+SEQ []"]
        L += ["=VAL 11a %.16g" % rng.uniform(-50, 50) for _ in range(6)]
        L += ["-SEQ", "-SEQ", "=VAL 11a sizes", "+SEQ []", "=VAL 11a %.16g" % rng.uniform(1, 4), "-SEQ", "-MAP"]
    L += ["-SEQ", "-MAP", "-DOC", "-STR"]
    listing = "\n".join(L) + "\n"
    rc, text = lite_emit(listing)
    assert rc == 0 and len(text) > 500000
    rc, parsed = lite_events(text)
    assert rc == 0
    # parsed listing carries the styles found (plain, flow/block); feeding it back writes the same text
    rc, text2 = lite_emit(parsed)
    assert rc == 0 and text2 == text


# ---- the reference's own scene.c emitter / parser running over yaml_lite -----------------------------

needs_ref = pytest.mark.skipif(not refharness.available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def ref():
    return refharness.RefHarness()


@needs_ref
def test_reference_emitter_over_yaml_lite_writes_what_libyaml_writes(ref, tmp_path):
    """scene_write_yaml (scene.c:1000) and scene_write_yaml_buffer (scene.c:1045, the MPI scene transport)
    for three scenes: the text equals libyaml's for the same event stream, and the committed fixture is
    what the reference writes today."""
    for scene, dims, frame, frames, cfg in (("mixed10d", 10, 0, 24, None), ("hypercube", 5, 3, 300, "hcube"),
                                            (None, 4, 0, 300, None)):
        ref.open_scene(scene)
        f1 = str(tmp_path / "a.yaml"); f2 = str(tmp_path / "b.yaml")
        ref.write_yaml(dims, frame, frames, f1, cfg)
        ref.write_yaml(dims, frame, frames, f2, cfg, to_buffer=True)
        text = open(f1, "rb").read()
        assert text == open(f2, "rb").read()
        err, listing = libyaml_events(text)
        assert err == 0
        rc, mine = lite_events(text)
        assert rc == 0 and mine == listing
        assert libyaml_emit(listing) == text
        if scene == "mixed10d":
            assert text == open(os.path.join(SCENES, "config5_mixed10d.yaml"), "rb").read()


@needs_ref
def test_reference_buffer_writer_grows_its_buffer(ref, tmp_path):
    """BASELINE config 2 as YAML is larger than scene_write_yaml_buffer's first 1 MiB buffer (scene.c:1062): the
    emitter must report the overflow the way libyaml's string writer does (size_written == size) so that the
    reference doubles the buffer and tries again (scene.c:1066-1085).  Both writers, one text, libyaml's events."""
    ref.open_scene("hypercube")
    f1 = str(tmp_path / "a.yaml"); f2 = str(tmp_path / "b.yaml")
    ref.write_yaml(8, 0, 300, f1, None)
    ref.write_yaml(8, 0, 300, f2, None, to_buffer=True)
    text = open(f1, "rb").read()
    assert len(text) > (1 << 20)
    assert text == open(f2, "rb").read()
    err, listing = libyaml_events(text)
    rc, mine = lite_events(text)
    assert err == 0 and rc == 0 and mine == listing
    assert listing.count("=VAL 10p type") >= 6561


@needs_ref
def test_yaml_scene_reload_is_a_fixed_point(ref, tmp_path):
    """C scene -> YAML -> scene -> YAML -> scene: the second and third generation are identical text and
    identical flat scenes.  (The first reload legitimately differs from the C scene: the ambient light
    becomes a list entry, scene.c:977-979 vs 1826-1829, and the reader has no `angle` key, scene.c:1772-1860.)"""
    import ndt_b200

    def load(path):
        ref.open_scene("yaml")
        assert ref.scene_frames(10, path) == 1
        ref.begin_frame(10, 0, 1, path)
        flat = ndt_b200.flatten(ref.scene_ptr, ref.kdtree_ptr, 64, 36, 128, 1, ref.get_bounds_ptr)
        ref.end_frame()
        return bytes(flat.blob)

    y1 = os.path.join(SCENES, "config5_mixed10d.yaml")
    y2 = str(tmp_path / "gen2.yaml"); y3 = str(tmp_path / "gen3.yaml")
    ref.open_scene("yaml"); ref.write_yaml(10, 0, 1, y2, y1)
    ref.open_scene("yaml"); ref.write_yaml(10, 0, 1, y3, y2)
    assert open(y2, "rb").read() == open(y3, "rb").read()
    a, b = load(y1), load(y2)
    assert a == b


@needs_ref
def test_yaml_frames_are_documents(ref):
    """scene_yaml_count_frames (scene.c:2134) and scene_yaml_skip_to_frame (scene.c:2067): frame k of the
    hand-written file is its k-th document (the sphere moves 5 units per frame)"""
    import ndt_b200
    cfg = "tests/scenes/handwritten4d.yaml"
    ref.open_scene("yaml")
    assert ref.scene_frames(4, cfg) == 3
    xs = []
    for f in range(3):
        ref.open_scene("yaml")
        ref.begin_frame(4, f, 3, cfg)
        flat = ndt_b200.flatten(ref.scene_ptr, ref.kdtree_ptr, 32, 18, 128, 1, ref.get_bounds_ptr)
        hit, oid, dist = ref.primary(32, 18)
        ref.end_frame()
        assert flat.header.n_objects == 7 and flat.header.n_items == 7     # the cluster is flattened (object.c:636-643)
        xs.append(np.frombuffer(bytes(flat.blob), np.float64))
    assert not np.array_equal(xs[0][: len(xs[1])], xs[1][: len(xs[0])])
