"""Teacher language channel: xw_sentence_compose against ALL sentences the reference's own CFG class
generates from the reference's own grammar text for the same bindings (tests/golden/sentences.json.gz, made by
tests/golden/gen_sentence_golden.py in the build container).  The reference draws productions from Python's
unseeded `random`, so parity is language equality, not a sentence-by-sentence match: every sentence we emit is in
the reference's set, and over many (env, episode) draws we emit every sentence of it."""
import ctypes as C
import gzip
import json
import os

import pytest

from xworld_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    with gzip.open(os.path.join(HERE, "golden", "sentences.json.gz"), "rt") as f:
        return json.load(f)


def compose(lib, buf, **kw):
    enc = lambda s: s.encode() if s is not None else None
    q = _abi.XwSentenceQuery(rules=kw["rules"], task=kw["task"], kind=kw["kind"], direction=kw.get("direction", 0),
                             name1=enc(kw.get("name1")), name2=enc(kw.get("name2")), color=enc(kw.get("color")),
                             seed=kw.get("seed", 1234), env_id=kw.get("env_id", 0), episode=kw.get("episode", 0), salt=kw.get("salt", 0))
    rc = lib.xw_sentence_compose(C.byref(q), buf, len(buf))
    return rc, buf.value.decode()


def test_language_equals_the_reference_grammar():
    lib = _abi.load()
    buf = C.create_string_buffer(256)
    entries = golden()
    assert len(entries) == 49
    for ent in entries:
        want = set(ent["sentences"])
        got = set()
        n = 40 * len(want) + 200
        for i in range(n):
            rc, s = compose(lib, buf, env_id=i % 997, episode=i // 997, salt=(i * 7) % 50, **{k: v for k, v in ent.items() if k != "sentences"})
            assert rc == len(s) and rc > 0, (ent, rc)
            assert s in want, (s, {k: v for k, v in ent.items() if k != "sentences"})
            got.add(s)
            if i > 4 * len(want) and got == want:
                break
        assert got == want, ("not generated", sorted(want - got)[:5], {k: v for k, v in ent.items() if k != "sentences"})


def test_sentence_is_a_pure_function_of_the_query_and_errors_are_statuses():
    lib = _abi.load()
    buf = C.create_string_buffer(256)
    kw = dict(rules=_abi.XW_RULES_NAV3D, task=3, kind=_abi.XW_SENT_START, name1="apple", direction=3, env_id=77, episode=5)
    a = compose(lib, buf, **kw)
    b = compose(lib, buf, **kw)
    assert a == b and "apple" in a[1] and "left" in a[1]
    assert compose(lib, buf, **dict(kw, env_id=78)) != a or compose(lib, buf, **dict(kw, env_id=79)) != a
    # walls.json: the two tasks that never leave idle in the reference commit have no command; no "wrong" sentence
    assert compose(lib, buf, rules=_abi.XW_RULES_NAV2D, task=1, kind=_abi.XW_SENT_START, name1="apple") == (0, "")
    assert compose(lib, buf, rules=_abi.XW_RULES_NAV2D, task=0, kind=_abi.XW_SENT_CORRECT)[1] == "Well done !"
    assert compose(lib, buf, rules=_abi.XW_RULES_NAV3D, task=0, kind=_abi.XW_SENT_TIMEUP)[1] == "Time up ."
    # statuses, not aborts
    assert compose(lib, buf, rules=_abi.XW_RULES_NAV3D, task=3, kind=_abi.XW_SENT_START, name1="apple", direction=0)[0] < 0
    assert compose(lib, buf, rules=_abi.XW_RULES_NAV3D, task=2, kind=_abi.XW_SENT_START, name1="apple")[0] < 0
    assert compose(lib, buf, rules=_abi.XW_RULES_NAV3D, task=9, kind=_abi.XW_SENT_START, name1="apple")[0] < 0
    small = C.create_string_buffer(4)
    assert compose(lib, small, **kw)[0] < 0
