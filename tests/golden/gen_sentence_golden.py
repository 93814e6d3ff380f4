#!/usr/bin/env python
"""Golden vectors for the teacher language channel (xw_sentence_compose): for every navigation task, sentence kind
and a sample of slot bindings, ALL sentences the REFERENCE's own CFG class generates from the REFERENCE's own grammar
text -- python/context_free_grammar.py (CFG.bind / CFG.generate_all) over the `grammar_str` of
games/xworld3d/tasks/XWorld3DNavTarget*.py and games/xworld/tasks/XWorldNav*.py, read from /root/reference at
generation time (nothing is copied into the repo but the resulting sentences).

Run in the build container (needs /root/reference):  python tests/golden/gen_sentence_golden.py
Writes tests/golden/sentences.json.gz."""
import gzip
import importlib.util
import itertools
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from xworld_b200.catalog import Catalog  # noqa: E402


class _Dict(dict):  # the reference is Python 2: CFG.__unbind_all calls dict.iteritems
    def iteritems(self):
        return self.items()


def load_cfg_class():
    spec = importlib.util.spec_from_file_location("ref_cfg", os.path.join(REF, "python", "context_free_grammar.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.CFG


def grammar_text(path):
    src = open(path).read()
    body = re.search(r'grammar_str = """(.*?)"""\s*%\s*(\(.*?\)|\w+)', src, re.S)
    return body.group(1), [a.strip() for a in body.group(2).strip("()").split(",")]


TASKS = [  # (rules, task id, file, slot nonterminals)
    (0, 0, "games/xworld3d/tasks/XWorld3DNavTarget.py"), (0, 1, "games/xworld3d/tasks/XWorld3DNavTargetNear.py"),
    (0, 2, "games/xworld3d/tasks/XWorld3DNavTargetBetween.py"), (0, 3, "games/xworld3d/tasks/XWorld3DNavTargetDirection.py"),
    (0, 4, "games/xworld3d/tasks/XWorld3DNavTargetAvoid.py"),
    (1, 0, "games/xworld/tasks/XWorldNavTarget.py"), (1, 2, "games/xworld/tasks/XWorldNavColorTarget.py"),
]
KINDS = {0: ["start", "correct", "wrong", "timeup"], 1: ["start", "finish", None, "timeup"]}
DIRS = {1: "FRONT", 2: "BEHIND", 3: "LEFT", 4: "RIGHT"}


def main():
    CFG = load_cfg_class()
    cat = Catalog.synthetic(seed=0)
    names = cat.names
    colors = sorted({m["color"] for m in cat.icon_meta if m["type"] == "goal" and m["color"] != "na"})
    rhs = lambda xs: "|".join("'" + x + "'" for x in xs)
    subst = {"all_goal_names": rhs(names), "all_colors": rhs(colors), "all_directions": rhs(["east", "west", "north", "south"])}
    sample_names = [names[0], names[len(names) // 2], names[-1]]
    out = []
    for rules, task, rel in TASKS:
        text, args = grammar_text(os.path.join(REF, rel))
        gstr = text % tuple(subst[a] for a in args)
        for kind, kname in enumerate(KINDS[rules]):
            if kname is None:
                continue
            if kind != 0:
                bindings = [dict()]
            elif rules == 0 and task == 2:
                bindings = [dict(name1=a, name2=b) for a, b in itertools.permutations(sample_names, 2)]
            elif rules == 0 and task == 3:
                bindings = [dict(name1=a, direction=d) for a in sample_names[:2] for d in DIRS]
            elif rules == 1 and task == 2:
                bindings = [dict(name1=a, color=c) for a in sample_names[:2] for c in colors[:2]]
            else:
                bindings = [dict(name1=a) for a in sample_names]
            for b in bindings:
                g = CFG(gstr, "S")
                g.productions = _Dict(g.productions)
                g.bind("S -> " + kname)
                if kind == 0:
                    if rules == 0 and task == 2:
                        g.bind("G1 -> '%s'" % b["name1"]); g.bind("G2 -> '%s'" % b["name2"])
                    elif rules == 1 and task == 2:
                        g.bind("O -> '%s'" % b["name1"]); g.bind("C -> '%s'" % b["color"])
                    else:
                        g.bind("G -> '%s'" % b["name1"])
                    if "direction" in b:
                        g.bind("P -> " + DIRS[b["direction"]])
                sentences = sorted(set(g.generate_all()))
                out.append(dict(rules=rules, task=task, kind=kind, sentences=sentences, **b))
    path = os.path.join(HERE, "sentences.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("%d bindings, %d sentences -> %s (%d bytes)" % (len(out), sum(len(o["sentences"]) for o in out), path, os.path.getsize(path)))


if __name__ == "__main__":
    main()
