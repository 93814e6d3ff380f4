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


def test_sentences_match_the_reference_python_sentence_by_sentence(synthetic_catalog):
    """The reference's own task code and CFG class, run with its `random` replayed from the same Philox substreams
    (tests/golden/gen_reference_python.py), against sentence_for_state over the oracle's state: the same sentence at
    reset and after every step -- which sentence is due (command / "Well done !" / nothing), its slots and its
    productions."""
    import ctypes as C
    import numpy as np
    import oracle
    from xworld_b200.simulator import sentence_for_state
    lib = _abi.load()
    olib = oracle.lib()
    with gzip.open(os.path.join(HERE, "golden", "refpy_traces.json.gz")) as f:
        tr = json.loads(f.read().decode())
    n_cmp = n_nonempty = 0
    kinds = set()
    for case in tr["cases"]:
        D, G = case["dim"], case["n_goals"]
        cfg = _abi.default_config(height=D, width=D, n_goals=G, n_blocks=case["n_blocks"], rules=case["rules"],
                                  seed=case["seed"], simulator_seed=case["simulator_seed"])
        for env in case["envs"]:
            cfg.env_id_offset = env["env_gid"]
            o = oracle.Oracle(cfg, synthetic_catalog, 1)
            e = o.envs[0]

            def current():
                st = dict(task=e.task, stage=e.stage, event=e.event, aux0=e.aux0, aux1=e.aux1, goal_name=list(e.goal_name),
                          goal_icon=list(e.goal_icon), episode=e.episode, steps_in_task=e.steps_in_task, num_steps=e.num_steps)
                return sentence_for_state(lib, cfg, synthetic_catalog, env["env_gid"], st)

            for ep in env["episodes"]:
                o.reset()
                where = (case["tag"], env["env_gid"], ep["episode"])
                assert current() == ep["reset_sentence"], (where, current(), ep["reset_sentence"])
                for i, s in enumerate(ep["steps"]):
                    r, ov = C.c_float(), C.c_int32()
                    assert olib.xo_step(C.byref(cfg), C.byref(o.cat_c), C.byref(e), s["a"], 1, C.byref(r), C.byref(ov)) == 0
                    got = current()
                    # (after the episode has ended the reference's terminal stage says nothing)
                    assert got == s["sent"], (where, i, got, s["sent"], s["stage"], s["ev"])
                    n_cmp += 1
                    n_nonempty += bool(s["sent"])
                    kinds.add((case["rules"], s["task"], s["sent"][:4]))
    assert n_cmp > 9000 and n_nonempty > 5000 and len(kinds) > 40
