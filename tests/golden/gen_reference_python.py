#!/usr/bin/env python
"""Golden-vector generator: runs the REFERENCE's own Python (map generator + teacher tasks) and
records what it does, so the C oracle (oracle/xw_oracle.c) can be pinned against it.

Runs only in the build container (needs /root/reference).  Output: tests/golden/refpy_traces.json.gz;
`gen_reference_python.py curriculum` writes tests/golden/refpy_curriculum.json.gz instead: XWorldNav with
--curriculum > 0 run unmodified at its own 8x8 / level tables, XWorldEnv.curriculum_check_period set
to a few resets so that the levels move within a short trace, plus one env at the reference's own period of 100.
`gen_reference_python.py fpv` writes tests/golden/refpy_fpv.json.gz: the same code with --visible_radius > 0:
set_property's agent yaw / goal yaw, scale, offset draws (xworld_env.py:207-223), the six first-person actions, and the
teacher's reach test with a heading that turns; and with --task_mode=one_channel (py_simulator.cpp:128-130, the reference's
Python default): the navigation2d.json tasks past their terminal stage, the walls.json tasks' time-up rule
(xworld_task.py:203-210; navigation group only -- the XWorldRec question tasks are outside the path).

What is executed unmodified (loaded from /root/reference, never copied):
    games/xworld/maps/xworld_env.py, XWorldNav.py      map/entity schema + generator
    python/maze2d.py, python/py_util.py                maze DFS, bfs, flood_fill
    games/xworld3d/tasks/xworld3d_task.py + XWorld3DNav{Target,TargetNear,TargetBetween,
        TargetDirection,TargetAvoid}.py                navigation2d.json task set
    python/context_free_grammar.py                     CFG: the teacher's sentences
    games/xworld/tasks/xworld_task.py + XWorldNav{Target,Near,ColorTarget,Between}.py   walls.json

What the harness supplies instead of the C++ host (and cites):
    * Python-2 -> 3 source fix-ups at load time: integer `/` on four lines (maze2d.py:89,
      xworld_env.py:129-130, xworld_task.py:206), dict.iteritems(), dict.keys() used as a list (xworld_env.py:292).
    * py_gflags.get_flag (python/py_init.cpp:37-58) -> a dict of the flags.
    * the `random` module -> ReplayRandom: every call site the reference draws from is mapped onto the
      oracle's Philox substream of the same name (oracle/xw_oracle.h XO_SITE_*), with the sequence
      put in a canonical order first wherever the reference's order is a CPython set/dict order.
    * XMap::move_item / XAgent::act (xmap.cpp:76-101, xitem.cpp:89-155) and Task::py_stage's
      C++<->Python sync (teaching_task.cpp:64-116) -> ~40 lines below (the compiled reference
      XMap/XAgent in oracle/_ref is checked against the same rules in tests/test_ref_lib.py).
    * XWorldNav's hard-coded 8x8 / goal / block counts (XWorldNav.py:10-11,31-32) are parametrised
      by text substitution for the 7x7 / 11x11 / 15x15 BASELINE configs; the 8x8 case runs as is.
"""
import json
import os
import re
import sys
import types

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import oracle  # noqa: E402  (only for xo_draw / xo_randbelow: the shared RNG definition)
from xworld_b200.catalog import Catalog  # noqa: E402

ITEM_PATH = os.path.join(REF, "games/xworld/images")

FLAGS = {"visible_radius": 0, "curriculum": 0, "task_mode": "lang_acquisition", "max_steps_factor": 10}


# ----------------------------------------------------------------------------- ReplayRandom
SITE_SENTENCE = 11  # xworld_b200/csrc/xw_sentence.hpp


class Ctx(object):
    sent_i = 0
    sent_salt = 0
    seed = 0
    env_gid = 0
    episode = 0
    attempt = 0
    maze_visits = 0
    goal_no = 0
    step_no = 0
    idle_calls = 0
    pose_calls = {}


def cell_key(p):  # canonical row-major (y, x)
    return (p[1], p[0])


class ReplayRandom(types.ModuleType):
    """Stands in for the `random` module inside the reference's Python."""

    def __init__(self):
        types.ModuleType.__init__(self, "random")

    def _u(self, site, index):
        return oracle.draw(Ctx.seed, Ctx.env_gid, Ctx.episode, Ctx.attempt, site, index)

    def _fy(self, x, site, base=0):  # python's shuffle: Fisher-Yates from the end
        for k, i in enumerate(reversed(range(1, len(x)))):
            j = oracle.randbelow(self._u(site, base + k), i + 1)
            x[i], x[j] = x[j], x[i]

    def shuffle(self, x):
        caller = sys._getframe(1).f_code.co_name
        if caller in ("__generate_all_grids", "_XWorldEnv__generate_all_grids", "update_entities_from_cpp", "bfs"):
            return  # result order is never observed (set() / canonicalised later / reachability only)
        if caller == "_configure":
            x.sort()
            self._fy(x, oracle.SITE_NAMES)
        elif caller == "dfs":
            self._fy(x, oracle.SITE_MAZE, base=3 * Ctx.maze_visits)
            Ctx.maze_visits += 1
        elif caller in ("__instantiate_entities", "_XWorldEnv__instantiate_entities"):
            self._fy(x, oracle.SITE_BLOCKS)  # blocks already in row-major generation order
        elif caller == "idle":
            if x and hasattr(x[0], "type"):  # random.shuffle(goals); g1, g2 = goals[:2]
                x.sort(key=lambda e: int(e.id.split("_")[-1]))
                self._fy(x, oracle.SITE_TASK_SHUF)
            else:  # random.shuffle(tiles); tiles[0]  == uniform pick (only element 0 is used)
                if x:
                    k = oracle.randbelow(self._u(oracle.SITE_TASK_A, 0), len(x))
                    x[0], x[k] = x[k], x[0]
        else:
            raise RuntimeError("unmapped shuffle call site: " + caller)

    def choice(self, seq):
        f1 = sys._getframe(1)
        caller = f1.f_code.co_name
        seq = list(seq)
        if caller == "check_or_get_value":
            ent = f1.f_back.f_locals["entity"]
            if isinstance(seq[0], tuple):  # loc = choice(available_grids)
                seq.sort(key=cell_key)
                if ent.type == "goal":
                    u = self._u(oracle.SITE_GOAL_LOC, Ctx.goal_no)
                else:
                    assert ent.type == "agent"
                    u = self._u(oracle.SITE_AGENT_LOC, 0)
                return seq[oracle.randbelow(u, len(seq))]
            if isinstance(seq[0], int):  # yaw = choice(range(-1, 3)) * PI_2 (xworld_env.py:208-210)
                assert ent.type == "agent" and seq == [-1, 0, 1, 2]
                return seq[oracle.randbelow(self._u(oracle.SITE_AGENT_YAW, 0), 4)]
            seq.sort()
            if len(seq) == 1:
                return seq[0]
            assert ent.type == "goal" and seq[0].endswith(".jpg")  # asset_path = choice(items[type][name])
            v = seq[oracle.randbelow(self._u(oracle.SITE_GOAL_ASSET, Ctx.goal_no), len(seq))]
            Ctx.goal_no += 1
            return v
        if caller == "value":  # RHS.value (context_free_grammar.py:41-49): the i-th production draw of this sentence
            i = Ctx.sent_i
            Ctx.sent_i += 1
            assert i < 16
            u = oracle.draw(Ctx.seed, Ctx.env_gid, Ctx.episode, 0, SITE_SENTENCE, ((Ctx.sent_salt & 0x3fff) << 4) + i)
            return seq[oracle.randbelow(u, len(seq))]
        if caller == "idle":
            cls = f1.f_locals["self"].__class__.__name__
            Ctx.idle_calls += 1
            if cls.startswith("XWorldNav"):  # 2-D tasks: one choice per idle stage, keyed by step
                return seq[oracle.randbelow(self._u(oracle.SITE_TASK_A, Ctx.step_no), len(seq))]
            if isinstance(seq[0], tuple) and isinstance(seq[0][0], tuple):  # choice(new_a): ((x,y,0), step)
                seq.sort(key=lambda t: cell_key(t[0]))  # BFS discovery order -> canonical row-major (the set is the contract)
                return seq[oracle.randbelow(self._u(oracle.SITE_TASK_AGENT, 0), len(seq))]
            if isinstance(seq[0], tuple):  # choice(empty_grids)
                seq.sort(key=cell_key)
                return seq[oracle.randbelow(self._u(oracle.SITE_TASK_B, 0), len(seq))]
            # entities: first choice = sel_goal (A), second = referent (B)
            site = oracle.SITE_TASK_A if Ctx.idle_calls == 1 else oracle.SITE_TASK_B
            seq.sort(key=lambda e: int(e.id.split("_")[-1]))
            return seq[oracle.randbelow(self._u(site, 0), len(seq))]
        raise RuntimeError("unmapped choice call site: " + caller)

    def uniform(self, a, b):
        """random.uniform(a, b) = a + (b - a) * random() (CPython random.py).  Only set_property draws continuous values:
        a goal's yaw, scale, offset, in that order (xworld_env.py:211-223) -> oracle site GOAL_POSE, index 4 * goal + k;
        random() = draw * 2^-32, and (draw >> 20) / 4096 for the yaw (oracle/xw_oracle_fpv.c xo_goal_pose)."""
        f1 = sys._getframe(1)
        assert f1.f_code.co_name == "check_or_get_value"
        ent = f1.f_back.f_locals["entity"]
        assert ent.type == "goal"
        g = int(ent.id.split("_")[-1])
        k = Ctx.pose_calls.get(g, 0)
        Ctx.pose_calls[g] = k + 1
        assert k < 3
        u = self._u(oracle.SITE_GOAL_POSE, 4 * g + k)
        r = (u >> 20) / 4096.0 if k == 0 else u * (1.0 / 4294967296.0)
        return a + (b - a) * r

    def randint(self, a, b):
        raise RuntimeError("unmapped randint()")

    def random(self):
        raise RuntimeError("unmapped random()")


# ----------------------------------------------------------------------------- module loading
def load_reference_python(dim, n_goals, n_blocks):
    """exec the reference sources (with the py2->3 fix-ups) into fresh modules."""
    rnd = ReplayRandom()
    mods = {}

    def mk(name, path, subs=()):
        src = open(os.path.join(REF, path)).read()
        for a, b in subs:
            assert re.search(a, src), (path, a)
            src = re.sub(a, b, src)
        m = types.ModuleType(name)
        m.__file__ = os.path.join(REF, path)
        sys.modules[name] = m
        mods[name] = m
        exec(compile(src, m.__file__, "exec"), m.__dict__)
        return m

    gf = types.ModuleType("py_gflags")
    gf.get_flag = lambda k: FLAGS[k]
    sys.modules["py_gflags"] = gf
    saved = sys.modules.get("random")
    sys.modules["random"] = rnd
    try:
        # the reference's own CFG class (RHS.value draws through ReplayRandom.choice, site SENTENCE)
        mk("context_free_grammar", "python/context_free_grammar.py", [(r"\.iteritems\(\)", ".items()")])
        mk("py_util", "python/py_util.py")
        mk("maze2d", "python/maze2d.py", [(r"\(X \+ 1\) / 2, \(Y \+ 1\) / 2", "(X + 1) // 2, (Y + 1) // 2")])
        mk("xworld_env", "games/xworld/maps/xworld_env.py", [
            (r"\(self\.max_height - h\) / 2", "(self.max_height - h) // 2"),
            (r"\(self\.max_width - w\) / 2", "(self.max_width - w) // 2"),
            (r"return self\.items\[type\]\.keys\(\)", "return list(self.items[type].keys())")])
        nav_subs = []
        if (dim, n_goals, n_blocks) != (8, 4, 16):
            nav_subs = [(r"max_height=8,", "max_height=%d," % dim), (r"max_width=8,", "max_width=%d," % dim),
                        (r"num_goals_seq = \[2, 2, 2, 4, 4, 4\]", "num_goals_seq = [%d] * n_levels" % n_goals),
                        (r"num_blocks_seq = \[0, 3, 6, 9, 12, 16\]", "num_blocks_seq = [%d] * n_levels" % n_blocks)]
        mk("XWorldNav", "games/xworld/maps/XWorldNav.py", nav_subs)
        mk("xworld3d_task", "games/xworld3d/tasks/xworld3d_task.py")
        for t in ("XWorld3DNavTarget", "XWorld3DNavTargetNear", "XWorld3DNavTargetBetween",
                  "XWorld3DNavTargetDirection", "XWorld3DNavTargetAvoid"):
            mk(t, "games/xworld3d/tasks/%s.py" % t)
        mk("xworld_task", "games/xworld/tasks/xworld_task.py", [(r"iteritems", "items"), (r"h\*w / 2", "h*w // 2")])
        for t in ("XWorldNavTarget", "XWorldNavNear", "XWorldNavColorTarget", "XWorldNavBetween"):
            mk(t, "games/xworld/tasks/%s.py" % t)
    finally:
        if saved is not None:
            sys.modules["random"] = saved
    return mods


T3 = ["XWorld3DNavTarget", "XWorld3DNavTargetNear", "XWorld3DNavTargetBetween", "XWorld3DNavTargetDirection",
      "XWorld3DNavTargetAvoid"]  # confs/navigation2d.json order
T2 = ["XWorldNavTarget", "XWorldNavNear", "XWorldNavColorTarget", "XWorldNavBetween"]  # confs/walls.json order
DIRS = {"front": 1, "behind": 2, "left": 3, "right": 4}


# ----------------------------------------------------------------------------- host emulation
class Host(object):
    """What XWorldSimulator/XWorld/XMap/Teacher do around the Python (C++ side), for one env."""
    start_level = 0

    def __init__(self, mods, cat, rules, dim):
        self.mods, self.cat, self.rules, self.dim = mods, cat, rules, dim
        self.env = mods["XWorldNav"].XWorldNav(ITEM_PATH, start_level=self.start_level)
        names = T3 if rules == 0 else T2
        self.tasks = [getattr(mods[n], n)(self.env) for n in names]  # Task::init_py_task
        self.path2icon = {os.path.join(ITEM_PATH, m["path"]): i for i, m in enumerate(cat.icon_meta)}
        self.stage = "idle"
        self.busy = None
        self.minstd = None

    # XWorld::reset (xworld.cpp:109-151): entity dicts in cpp_get_entities order
    def pull_entities(self):
        self.entities = [dict(e) for e in self.env.cpp_get_entities()]
        for e in self.entities:
            e["loc"] = tuple(float(v) for v in e["loc"])  # Entity ctor: doubles (simulator_entity.h:88-101)

    def cell(self, x, y):
        for e in self.entities:
            if int(e["loc"][0]) == x and int(e["loc"][1]) == y:
                return e
        return None

    def agent(self):
        return [e for e in self.entities if e["type"] == "agent"][0]

    @staticmethod
    def facing(yaw):  # XItem::get_item_facing_dir (xitem.cpp:65-78): 0 right, 1 down, 2 left, 3 up
        import math
        return 0 if abs(yaw) < 1e-4 else 1 if abs(yaw - math.pi / 2) < 1e-4 else 2 if abs(yaw - math.pi) < 1e-4 else 3

    # XAgent::act + XMap::move_item
    def move(self, action):
        import math
        a = self.agent()
        if FLAGS["visible_radius"]:  # xitem.cpp:100-151; the compiled XAgent is checked against the oracle in tests/test_oracle_fpv.py
            fx, fy = [(1, 0), (0, 1), (-1, 0), (0, -1)][self.facing(a["yaw"])]
            if action >= 4:
                if action == 4:
                    a["yaw"] -= math.pi / 2
                    if a["yaw"] < -math.pi / 2 - 1e-4:
                        a["yaw"] += 2 * math.pi
                else:
                    a["yaw"] += math.pi / 2
                    if a["yaw"] > math.pi + 1e-4:
                        a["yaw"] -= 2 * math.pi
                return False, []  # move_item onto the agent's own cell: not reachable, no contact (xmap.cpp:76-101)
            dx, dy = [(fx, fy), (-fx, -fy), (fy, -fx), (-fy, fx)][action]
        else:
            dx, dy = [(0, -1), (0, 1), (-1, 0), (1, 0)][action]
        tx, ty = int(a["loc"][0]) + dx, int(a["loc"][1]) + dy
        contacts = []
        ok = False
        if 0 <= tx < self.dim and 0 <= ty < self.dim:
            it = self.cell(tx, ty)
            if it is None:
                a["loc"] = (float(tx), float(ty), 0.0)
                ok = True
            elif it["id"] != a["id"]:
                contacts.append(it["id"])
        return ok, contacts

    # Task::py_stage (teaching_task.cpp:64-116)
    def py_stage(self, task, stage, success, game_event):
        env = self.env
        env.update_entities_from_cpp([dict(e) for e in self.entities])
        env.update_agent_sentence_from_cpp("")
        env.update_agent_action_success_from_cpp(success)
        env.update_game_event_from_cpp(game_event)
        Ctx.idle_calls = 0
        Ctx.sent_i = 0
        Ctx.sent_salt = 0 if self.rules == 0 else Ctx.step_no  # walls.json tasks issue several commands per episode
        ret = getattr(task, stage)()
        if env.env_changed():
            self.pull_entities()  # XWorldSimulator::update_environment -> XWorld::reset(false)
        event = task.get_event()
        assert len(ret) == 3
        self.last_sentence = ret[2]
        return ret[0], float(ret[1]), event

    def snapshot(self):
        D = self.dim
        grid = [[0] * D for _ in range(D)]
        goals = sorted([e for e in self.entities if e["type"] == "goal"], key=lambda e: int(e["id"].split("_")[-1]))
        for e in self.entities:
            x, y = int(e["loc"][0]), int(e["loc"][1])
            if e["type"] == "block":
                grid[y][x] = 1
            elif e["type"] == "agent":
                grid[y][x] = 2
            else:
                grid[y][x] = 3 + goals.index(e)
        a = self.agent()
        return {
            "grid": [v for row in grid for v in row],
            "agent": [int(a["loc"][0]), int(a["loc"][1])],
            "goal_x": [int(g["loc"][0]) for g in goals], "goal_y": [int(g["loc"][1]) for g in goals],
            "goal_name": [self.cat.names.index(g["name"]) for g in goals],
            "goal_icon": [self.path2icon[g["asset_path"]] for g in goals],
            "agent_yaw": a["yaw"],
            "goal_pose": [[g["yaw"], g["scale"], g["offset"]] for g in goals],
        }

    def approach_from_above(self, goal):
        """First move of a shortest path to the cell above `goal`, MOVE_DOWN once there (the only scoring move: the
        heading is +y); None when there is no path.  Harness-side action policy only."""
        occ = {(int(e["loc"][0]), int(e["loc"][1])) for e in self.entities}
        a = self.agent()
        start = (int(a["loc"][0]), int(a["loc"][1]))
        dest = (int(goal["loc"][0]), int(goal["loc"][1]) - 1)
        if start == dest:
            return 1
        if dest in occ or dest[1] < 0:
            return None
        prev = {start: None}
        queue = [start]
        moves = [(0, -1, 0), (0, 1, 1), (-1, 0, 2), (1, 0, 3)]
        while queue:
            cur = queue.pop(0)
            if cur == dest:
                while prev[cur][0] != start:
                    cur = prev[cur][0]
                return prev[cur][1]
            for dx, dy, act in moves:
                q = (cur[0] + dx, cur[1] + dy)
                if 0 <= q[0] < self.dim and 0 <= q[1] < self.dim and q not in occ and q not in prev:
                    prev[q] = (cur, act)
                    queue.append(q)
        return None

    def goal_index(self, ent):
        return int(ent.id.split("_")[-1])

    def task_record(self, t, task):
        """target bookkeeping in the oracle's encoding (xo_env.target_mask / aux*)."""
        rec = {}
        if self.rules == 0:
            name = T3[t]
            if name in ("XWorld3DNavTarget", "XWorld3DNavTargetAvoid", "XWorld3DNavTargetNear"):
                rec["target_mask"] = sum(1 << self.goal_index(g) for g in task.target)
            elif name == "XWorld3DNavTargetBetween":
                o1, o2 = task.target  # the Python side's coordinates; the trace keeps map coordinates (cpp_get_entities)
                rec["mid"] = [int((o1[0] + o2[0]) // 2) + self.env.offset_w, int((o1[1] + o2[1]) // 2) + self.env.offset_h]
            else:
                referent, direction = task.target
                rec["referent"] = self.goal_index(referent)
                rec["direction"] = DIRS[direction]
        return rec


def run_case(mods, cat, rules, dim, n_goals, n_blocks, seed, simulator_seed, env_gid, n_episodes, n_steps, act_seed,
             curriculum=False):
    """One env, several episodes; returns the trace the oracle must reproduce."""
    L = oracle.lib()
    import ctypes as C
    host = Host(mods, cat, rules, dim)
    minstd = C.c_uint32(L.xo_minstd_seed_for_thread(simulator_seed, env_gid + 1))
    Ctx.seed, Ctx.env_gid = seed, env_gid
    episodes = []
    import numpy as np
    arng = np.random.RandomState(act_seed)
    for ep in range(1, n_episodes + 1):
        Ctx.episode = ep
        rec = {"episode": ep}
        # ---- SimulatorInterface::reset_game
        if rules == 0:
            t = L.xo_get_rand_ind(C.byref(minstd), 5)  # TaskGroup::run_stage (seeded C++ engine)
            task = host.tasks[t]
            ok = False
            for att in range(64):
                Ctx.attempt, Ctx.maze_visits, Ctx.goal_no, Ctx.step_no = att, 0, 0, 0
                Ctx.pose_calls = {}
                if att > 0:  # a re-drawn map is the same reset: the curriculum check ran with attempt 0
                    host.env.get_current_usage = lambda: 0
                host.env.reset()
                host.env.__dict__.pop("get_current_usage", None)
                assert host.env.env_changed()
                host.pull_entities()
                task.reset()
                try:
                    stage, r0, ev0 = host.py_stage(task, "idle", False, "")
                    ok = True
                    break
                except AssertionError as ex:
                    if "crowded" not in str(ex):
                        raise
            assert ok
            rec["attempts"] = att + 1
            assert stage == "navigation_reward" and r0 == 0.0
        else:
            Ctx.attempt, Ctx.maze_visits, Ctx.goal_no, Ctx.step_no = 0, 0, 0, 0
            Ctx.pose_calls = {}
            host.env.reset()
            host.pull_entities()
            stage, task, t = "idle", None, -1
        rec["task"] = t
        rec["reset"] = host.snapshot()
        if curriculum:
            rec["level"] = host.env.current_level
            rec["dims"] = list(host.env.get_dims())
            rec["check_counter"] = host.env.curriculum_check_counter
            rec["usage"] = {k: [len(v), sum(v)] for k, v in host.env.current_usage.items()}
        if rules == 0:
            rec["reset_sentence"] = host.last_sentence
        if rules == 0:
            rec.update(host.task_record(t, task))
        steps = []
        num_steps = 0

        def teach2d(success, game_event):
            nonlocal stage, task, t
            if stage == "idle":
                t = L.xo_get_rand_ind(C.byref(minstd), 4)
                task = host.tasks[t]
                task.reset()
                stage, r, ev = host.py_stage(task, "idle", success, game_event)
            else:
                stage, r, ev = host.py_stage(task, stage, success, game_event)
            # XWorldRec group: one engine draw per teach, event overwritten with ""
            minstd.value = (minstd.value * 16807) % 2147483647
            return r, ""

        if rules == 1:
            r, ev = teach2d(False, "")
            rec["reset_stage"] = stage
            rec["reset_task"] = t
            rec["reset_sentence"] = host.last_sentence
        for s in range(n_steps):
            a = int(arng.randint(0, 6 if FLAGS["visible_radius"] else 4))
            if FLAGS["visible_radius"] and arng.rand() < 0.35:
                a = 0  # MOVE_FORWARD is the only move that can score: lets episodes end by reaching goals
            if curriculum and arng.rand() < 0.4:
                a = 1  # MOVE_DOWN is the only move that can score (the heading is +y): lets the success rates rise
            if curriculum and rules == 0 and T3[t] == "XWorld3DNavTargetDirection" and arng.rand() < 0.7:
                # walk to the goal next to the referent and bump it from above, so that the displaced-referent test
                # of the padded levels (oracle/xw_oracle.c xo_teach) sees goals reached in every arrangement
                ref_id = task.target[0].id
                goals = [e for e in host.entities if e["type"] == "goal"]
                ref = [e for e in goals if e["id"] == ref_id][0]
                near = [e for e in goals if e["id"] != ref_id and
                        abs(e["loc"][0] - ref["loc"][0]) + abs(e["loc"][1] - ref["loc"][1]) == 1]
                if near:
                    b = host.approach_from_above(near[0])
                    if b is not None:
                        a = b
            num_steps += 1
            Ctx.step_no = num_steps
            ok, contacts = host.move(a)
            game_event = ("collision:" + "|".join(contacts) + "\n") if contacts else ""
            if rules == 0:
                stage, r, ev = host.py_stage(task, stage, ok, game_event)
            else:
                r, ev = teach2d(ok, game_event)
            ag = host.agent()
            if curriculum:
                steps.append({"a": a, "ok": int(ok), "r": r, "ev": ev, "agent": [int(ag["loc"][0]), int(ag["loc"][1])]})
            else:
                steps.append({"a": a, "ok": int(ok), "r": r, "ev": ev, "stage": stage, "task": t, "sent": host.last_sentence,
                              "agent": [int(ag["loc"][0]), int(ag["loc"][1])], "yaw": ag["yaw"]})
            if curriculum and stage == "terminal":
                break  # the trainer resets a finished game (test_xworld.py:46-49)
            if rules == 0 and stage == "terminal" and s + 3 < n_steps and len(steps) > 2 and steps[-2]["stage"] == "terminal" \
                    and steps[-3]["stage"] == "terminal" and (FLAGS["task_mode"] != "one_channel" or len(steps) > 12):
                break  # a few terminal-stage steps are enough
        rec["steps"] = steps
        rec["minstd"] = minstd.value
        episodes.append(rec)
    return episodes


def main_curriculum():
    """XWorldNav.py:36-58 + xworld_env.py:103-110,118-134,352-366,454-473 + the task classes' success windows."""
    import gzip
    import numpy as np
    cat = Catalog.from_item_path(ITEM_PATH)
    thr = float(np.float32(0.05))  # FLAGS_curriculum = the option read as a float (py_simulator.cpp:127)
    FLAGS["curriculum"] = thr
    out = {"generator": "tests/golden/gen_reference_python.py curriculum", "curriculum": thr, "cases": []}
    cases = [("period5", 5, 24, 30, 60, 3, 0), ("period100_reference", 100, 1, 330, 12, 1, 0)]
    cases += [("start_level%d" % k, 4, 6 if k in (2, 3) else 4, 24 if k in (2, 3) else 14, 60, 1, k) for k in range(1, 6)]  # XWorldNav(item_path, start_level=k)
    # Envs in which the displaced-referent rule of the padded levels lets a Direction episode SUCCEED: found
    # by running this same action policy against the oracle (gids 0..1500), then replayed here by the reference itself.
    cases += [("level2_direction_success", 4, [405, 249], 10, 60, 1, 2), ("level3_direction_success", 4, [68, 206], 4, 60, 1, 3)]
    cases = [c + (0,) for c in cases]
    # walls.json rules (test_xworld.py:31-40 runs them with curriculum 0.1): in lang_acquisition mode no task class ever
    # records a result (xworld_task.py:205-214: time-up is one_channel only, a goal cell is never entered; the
    # XWorldRec* tasks stay in idle), so usage stays empty and the env keeps its start level: a padded world for ever
    cases += [("nav2d_level0", 4, 3, 8, 30, 10, 0, 1), ("nav2d_start_level3", 4, 3, 6, 40, 10, 3, 1)]
    for tag, period, n_envs, n_ep, n_st, msf, start_level, rules in cases:
        FLAGS["max_steps_factor"] = msf
        Host.start_level = start_level
        mods = load_reference_python(8, 4, 16)  # XWorldNav.py as it is
        mods["xworld_env"].XWorldEnv.curriculum_check_period = period  # a class attribute (xworld_env.py:58)
        envs = []
        for gid in (range(n_envs) if isinstance(n_envs, int) else n_envs):
            envs.append({"env_gid": gid,
                         "episodes": run_case(mods, cat, rules, 8, 4, 16, seed=4321, simulator_seed=3, env_gid=gid,
                                              n_episodes=n_ep, n_steps=n_st, act_seed=2000 + gid, curriculum=True)})
        levels = [ep["level"] for e in envs for ep in e["episodes"]]
        print(tag, "levels reached:", sorted(set(levels)))
        out["cases"].append({"tag": tag, "rules": rules, "dim": 8, "n_goals": 4, "n_blocks": 16, "seed": 4321, "simulator_seed": 3,
                             "check_period": period, "max_steps_factor": msf, "start_level": start_level, "envs": envs})
    path = os.path.join(HERE, "refpy_curriculum.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print("wrote", path, os.path.getsize(path))


def main_fpv():
    """--visible_radius > 0 and --task_mode=one_channel traces -> refpy_fpv.json.gz."""
    import gzip
    cat = Catalog.from_item_path(ITEM_PATH)
    cases = [  # (tag, rules, dim, goals, blocks, visible_radius, task_mode, n_envs, episodes, steps)
        ("fpv_nav3d_8x8_vr3", 0, 8, 4, 16, 3, "lang_acquisition", 8, 6, 90),
        ("fpv_nav3d_7x7_vr7", 0, 7, 4, 12, 7, "lang_acquisition", 8, 6, 90),
        ("fpv_nav3d_8x8_vr5_one_channel", 0, 8, 4, 16, 5, "one_channel", 4, 4, 60),
        ("fpv_nav2d_11x11_vr7", 1, 11, 4, 30, 7, "lang_acquisition", 6, 3, 70),
        ("nav3d_8x8_one_channel", 0, 8, 4, 16, 0, "one_channel", 6, 4, 60),
        ("nav2d_11x11_one_channel_nav_group", 1, 11, 4, 30, 0, "one_channel", 4, 2, 150),
        ("nav2d_8x8_one_channel_nav_group", 1, 8, 4, 16, 0, "one_channel", 4, 2, 100),
    ]
    out = {"generator": "tests/golden/gen_reference_python.py fpv", "cases": []}
    for tag, rules, dim, G, B, vr, mode, n_envs, n_ep, n_st in cases:
        FLAGS["visible_radius"], FLAGS["task_mode"] = vr, mode
        mods = load_reference_python(dim, G, B)
        envs = []
        for gid in range(n_envs):
            envs.append({"env_gid": gid,
                         "episodes": run_case(mods, cat, rules, dim, G, B, seed=777, simulator_seed=5,
                                              env_gid=gid, n_episodes=n_ep, n_steps=n_st, act_seed=3000 + gid)})
        evs = {}
        for e in envs:
            for ep in e["episodes"]:
                for st in ep["steps"]:
                    evs[st["ev"]] = evs.get(st["ev"], 0) + 1
        print(tag, "events:", evs)
        out["cases"].append({"tag": tag, "rules": rules, "dim": dim, "n_goals": G, "n_blocks": B, "seed": 777,
                             "simulator_seed": 5, "visible_radius": vr, "task_mode": mode, "envs": envs})
    FLAGS["visible_radius"], FLAGS["task_mode"] = 0, "lang_acquisition"
    path = os.path.join(HERE, "refpy_fpv.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print("wrote", path, os.path.getsize(path))


def main():
    cat = Catalog.from_item_path(ITEM_PATH)
    meta_names = Catalog(Catalog.metadata(), __import__("numpy").zeros((363, 64, 64, 3), "uint8")).names
    assert meta_names == cat.names
    cases = [  # (tag, rules, dim, goals, blocks, n_envs, episodes, steps)
        ("nav3d_8x8_reference_defaults", 0, 8, 4, 16, 6, 6, 80),
        ("nav3d_7x7", 0, 7, 4, 12, 10, 8, 80),
        ("nav3d_15x15", 0, 15, 4, 56, 3, 3, 60),
        ("nav2d_11x11", 1, 11, 4, 30, 6, 3, 60),
        ("nav2d_8x8_reference_defaults", 1, 8, 4, 16, 4, 3, 40),
    ]
    out = {"generator": "tests/golden/gen_reference_python.py", "cases": []}
    for tag, rules, dim, G, B, n_envs, n_ep, n_st in cases:
        mods = load_reference_python(dim, G, B)
        envs = []
        for gid in range(n_envs):
            envs.append({"env_gid": gid,
                         "episodes": run_case(mods, cat, rules, dim, G, B, seed=1234, simulator_seed=1,
                                              env_gid=gid, n_episodes=n_ep, n_steps=n_st, act_seed=1000 + gid)})
        out["cases"].append({"tag": tag, "rules": rules, "dim": dim, "n_goals": G, "n_blocks": B, "seed": 1234,
                             "simulator_seed": 1, "envs": envs})
        print(tag, "ok")
    import gzip
    path = os.path.join(HERE, "refpy_traces.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    if sys.argv[1:] == ["curriculum"]:
        main_curriculum()
    elif sys.argv[1:] == ["fpv"]:
        main_fpv()
    else:
        main()
