"""Pins the CPU oracle (oracle/xw_oracle.c) against everything the reference offers for this path:
its own golden vectors (tests/test_simulator_seed.cpp, tests/test_simple_game_simulator.cpp), its
own Python executed with a replayed RNG (tests/golden/refpy_traces.json.gz), its own C++ compiled
against header shims (oracle/_ref), and the published Philox known-answer vectors."""
import ctypes as C
import gzip
import json
import os

import numpy as np
import pytest

import oracle
from xworld_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))


def test_philox_known_answers(oracle_lib):
    # Random123 kat_vectors, philox4x32-10
    kats = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
            ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
            ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
             [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kats:
        out = (C.c_uint32 * 4)()
        oracle_lib.xo_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert list(out) == want


def test_reference_seed_golden_vectors(oracle_lib):
    """tests/test_simulator_seed.cpp:23-25: seed 1 on threads 1..5, then seed 2 on threads 6..10
    (the thread counter is a process-wide static, simulator_util.cpp:36-47)."""
    def seq(seed, threads):
        out = []
        for t in threads:
            s = C.c_uint32(oracle_lib.xo_minstd_seed_for_thread(seed, t))
            out.append(oracle_lib.xo_get_rand_ind(C.byref(s), 1000000))
        return out
    assert seq(1, range(1, 6)) == [266148, 605992, 817626, 635637, 393423]
    assert seq(2, range(6, 11)) == [258945, 847424, 238883, 918571, 875562]


def test_simple_game_known_answer(oracle_lib):
    """tests/test_simple_game_simulator.cpp:21-47: array_size 8, action 1 x3 -> -0.1, -0.1, 2.0."""
    g = oracle.XoSimpleGame()
    oracle_lib.xo_sg_reset(C.byref(g), 8)
    assert list(g.state)[:8] == [0, 0, 0, 0, 1, 0, 0, 0]
    rewards = [oracle_lib.xo_sg_act(C.byref(g), 1) for _ in range(3)]
    assert abs(rewards[0] + 0.1) < 1e-6 and abs(rewards[1] + 0.1) < 1e-6 and abs(rewards[2] - 2.0) < 1e-6
    assert oracle_lib.xo_sg_game_over(C.byref(g)) == 1


EVMAP = {"": 0, "correct_goal": 1, "wrong_goal": 2, "time_up": 3}


def test_against_reference_python_traces(oracle_lib, synthetic_catalog):
    """Map generation, the 5 XWorld3DNav* + 4 XWorldNav* idle stages and every per-step reward /
    event, as produced by the reference's own Python (gen_reference_python.py)."""
    with gzip.open(os.path.join(HERE, "golden", "refpy_traces.json.gz")) as f:
        tr = json.loads(f.read().decode())
    n_steps = n_resets = 0
    for case in tr["cases"]:
        D = case["dim"]
        cfg = _abi.default_config(height=D, width=D, n_goals=case["n_goals"], n_blocks=case["n_blocks"],
                                  rules=case["rules"], seed=case["seed"], simulator_seed=case["simulator_seed"])
        for env in case["envs"]:
            cfg.env_id_offset = env["env_gid"]
            o = oracle.Oracle(cfg, synthetic_catalog, 1)
            e = o.envs[0]
            for ep in env["episodes"]:
                o.reset()
                n_resets += 1
                rs = ep["reset"]
                where = (case["tag"], env["env_gid"], ep["episode"])
                assert list(e.grid)[:D * D] == rs["grid"], where
                assert [e.agent_x, e.agent_y] == rs["agent"], where
                G = case["n_goals"]
                assert list(e.goal_x)[:G] == rs["goal_x"] and list(e.goal_y)[:G] == rs["goal_y"], where
                assert list(e.goal_name)[:G] == rs["goal_name"] and list(e.goal_icon)[:G] == rs["goal_icon"], where
                if case["rules"] == 0:
                    assert e.task == ep["task"], where
                    if "target_mask" in ep:
                        assert e.target_mask == ep["target_mask"], where
                    if "mid" in ep:
                        assert [e.aux1, e.aux2] == ep["mid"], where
                    if "referent" in ep:
                        assert (e.aux0, e.aux1) == (ep["referent"], ep["direction"]), where
                else:
                    assert e.task == ep["reset_task"] and (e.stage != 0) == (ep["reset_stage"] != "idle"), where
                for i, s in enumerate(ep["steps"]):
                    r, ov = C.c_float(), C.c_int32()
                    rc = oracle_lib.xo_step(C.byref(cfg), C.byref(o.cat_c), C.byref(e), s["a"], 1, C.byref(r), C.byref(ov))
                    assert rc == 0
                    n_steps += 1
                    assert np.float32(r.value).view(np.uint32) == np.float32(s["r"]).view(np.uint32), (where, i)
                    assert [e.agent_x, e.agent_y] == s["agent"] and e.action_success == s["ok"], (where, i)
                    assert e.event == EVMAP[s["ev"]], (where, i)
                    # XWorldSimulator::game_over, lang_acquisition (xworld_simulator.cpp:165-177); --max_steps is off here
                    assert ov.value == {"": 0, "correct_goal": _abi.XW_SUCCESS, "wrong_goal": _abi.XW_DEAD,
                                        "time_up": _abi.XW_MAX_STEP}[s["ev"]], (where, i)
                assert e.minstd == ep["minstd"], where
    assert n_steps > 9000 and n_resets > 80


def test_curriculum_against_reference_python(oracle_lib, synthetic_catalog):
    """XWorldNav with --curriculum > 0, run by the reference's own Python (XWorldNav.py:36-58,
    xworld_env.py:103-110,118-134,352-366,454-473; the task classes' 200-result windows, xworld3d_task.py:129-146):
    the level of every episode, the padded 8x8 map, the counters and windows at every reset, and every step."""
    with gzip.open(os.path.join(HERE, "golden", "refpy_curriculum.json.gz")) as f:
        tr = json.loads(f.read().decode())
    T3 = ["XWorld3DNavTarget", "XWorld3DNavTargetNear", "XWorld3DNavTargetBetween", "XWorld3DNavTargetDirection",
          "XWorld3DNavTargetAvoid"]
    n_steps = n_resets = 0
    levels, ups = set(), 0
    displaced_ok = 0  # Direction episodes answered correctly at a padded level with offset 1
    for case in tr["cases"]:
        cfg = _abi.default_config(height=8, width=8, n_goals=4, n_blocks=16, rules=case["rules"], seed=case["seed"],
                                  simulator_seed=case["simulator_seed"], curriculum=tr["curriculum"],
                                  curriculum_check_period=case["check_period"], max_steps_factor=case["max_steps_factor"],
                                  start_level=case["start_level"])
        for env in case["envs"]:
            cfg.env_id_offset = env["env_gid"]
            o = oracle.Oracle(cfg, synthetic_catalog, 1)
            e = o.envs[0]
            prev = case["start_level"]
            for ep in env["episodes"]:
                o.reset()
                n_resets += 1
                rs = ep["reset"]
                where = (case["tag"], env["env_gid"], ep["episode"])
                assert e.level == ep["level"] and [e.dim, e.dim] == ep["dims"], where
                levels.add(e.level)
                ups += e.level != prev
                prev = e.level
                assert e.check_counter == ep["check_counter"], where
                if case["rules"] == 1:  # walls.json rules: nothing is ever recorded, the level stays
                    assert ep["usage"] == {} and e.level == case["start_level"] and ep["check_counter"] == ep["episode"], where
                    assert list(e.seq_len) == [0] * 5, where
                for t, name in enumerate(T3):
                    ln, sm = ep["usage"].get(name, [0, 0])
                    assert e.seq_len[t] == ln and sum(e.seq[t][:ln]) == sm, (where, name)
                G = len(rs["goal_x"])
                assert e.n_goals == G == (2 if e.level < 3 else 4), where
                assert list(e.grid)[:64] == rs["grid"], where
                assert [e.agent_x, e.agent_y] == rs["agent"], where
                assert list(e.goal_x)[:G] == rs["goal_x"] and list(e.goal_y)[:G] == rs["goal_y"], where
                assert list(e.goal_name)[:G] == rs["goal_name"] and list(e.goal_icon)[:G] == rs["goal_icon"], where
                if case["rules"] == 1:
                    assert e.task == ep["reset_task"] and (e.stage != 0) == (ep["reset_stage"] != "idle"), where
                else:
                    assert e.task == ep["task"], where
                if "target_mask" in ep:
                    assert e.target_mask == ep["target_mask"], where
                if "mid" in ep:
                    assert [e.aux1, e.aux2] == ep["mid"], where
                if "referent" in ep:
                    assert (e.aux0, e.aux1) == (ep["referent"], ep["direction"]), where
                for i, s in enumerate(ep["steps"]):
                    r, ov = C.c_float(), C.c_int32()
                    rc = oracle_lib.xo_step(C.byref(cfg), C.byref(o.cat_c), C.byref(e), s["a"], 1, C.byref(r), C.byref(ov))
                    assert rc == 0
                    n_steps += 1
                    assert np.float32(r.value).view(np.uint32) == np.float32(s["r"]).view(np.uint32), (where, i)
                    assert [e.agent_x, e.agent_y] == s["agent"] and e.action_success == s["ok"], (where, i)
                    assert e.event == EVMAP[s["ev"]], (where, i)
                assert e.minstd == ep["minstd"], where
                displaced_ok += (case["rules"] == 0 and ep["task"] == 3 and ep["level"] in (2, 3) and
                                 ep["steps"][-1]["ev"] == "correct_goal")
    assert displaced_ok >= 4
    assert levels == {0, 1, 2, 3, 4, 5} and ups >= 2 and n_resets > 1300 and n_steps > 10000, (levels, ups, n_resets, n_steps)


def test_reward_bit_patterns(oracle_lib, synthetic_catalog):
    """the float32 images of the reference's double sums."""
    want3 = {0xbc23d70a, 0x3f7d70a4, 0xbf8147ae, 0x00000000}
    want2 = {0xbdcccccd, 0xbe99999a, 0x00000000}
    for rules, want in ((0, want3), (1, want2)):
        cfg = _abi.default_config(height=7, width=7, n_goals=4, n_blocks=12, rules=rules, seed=5, simulator_seed=3)
        o = oracle.Oracle(cfg, synthetic_catalog, 64, threads=2)
        o.reset()
        seen = set()
        rng = np.random.RandomState(1)
        for _ in range(200):
            r, ov, _f = o.step(rng.randint(0, 4, 64))
            seen |= set(int(x) for x in r.view(np.uint32))
        assert seen <= want and len(seen) >= len(want) - 1, [hex(x) for x in seen]


def test_time_up(oracle_lib, synthetic_catalog):
    """_time_reward (xworld3d_task.py:472-482): steps >= h*w*max_steps_factor -> time_up -> MAX_STEP."""
    cfg = _abi.default_config(height=7, width=7, n_goals=4, n_blocks=12, rules=0, seed=9, simulator_seed=2,
                              max_steps_factor=1)
    o = oracle.Oracle(cfg, synthetic_catalog, 32)
    o.reset()
    hit = 0
    for s in range(49):
        r, ov, _f = o.step(np.zeros(32, np.int32))  # MOVE_UP never reaches a goal (heading is "down")
        if s < 48:
            assert (ov == 0).all()
    assert (ov == _abi.XW_MAX_STEP).all() and (r.view(np.uint32) == 0xbc23d70a).all()
    r, ov, _f = o.step(np.zeros(32, np.int32))
    assert (ov == 0).all() and (r == 0).all()  # terminal stage: reward 0, event gone after one step


# ---------------------------------------------------------------------------- oracle/_ref
def _ref():
    if not os.path.exists(oracle.REF_LIB):
        pytest.skip("oracle/_ref/libxw_ref.so not built (needs /root/reference)")
    R = C.CDLL(oracle.REF_LIB)
    R.ref_sg_create.restype = C.c_void_p
    R.ref_sg_take_action.restype = C.c_float
    R.ref_sg_take_action.argtypes = [C.c_void_p, C.c_int]
    R.ref_sg_game_over.argtypes = [C.c_void_p]
    R.ref_sg_screen.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    R.ref_sg_reset.argtypes = [C.c_void_p]
    R.ref_race_create.restype = C.c_void_p
    R.ref_race_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int]
    R.ref_race_step.restype = C.c_float
    R.ref_race_step.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    R.ref_race_reset.argtypes = [C.c_void_p]
    R.ref_map_create.restype = C.c_void_p
    R.ref_map_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int]
    R.ref_map_act.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
    R.ref_map_destroy.argtypes = [C.c_void_p]
    return R


def test_compiled_reference_reproduces_its_own_vectors():
    R = _ref()
    out = (C.c_int * 5)()
    R.ref_rand_ind_threads(1, 5, 1000000, out)
    assert list(out) == [266148, 605992, 817626, 635637, 393423]
    R.ref_rand_ind_threads(2, 5, 1000000, out)
    assert list(out) == [258945, 847424, 238883, 918571, 875562]
    g = R.ref_sg_create(8)
    rs = [R.ref_sg_take_action(g, 1) for _ in range(3)]
    assert abs(rs[0] + 0.1) < 1e-6 and abs(rs[2] - 2.0) < 1e-6 and R.ref_sg_game_over(g) == 4


def test_simple_game_port_vs_compiled_reference(oracle_lib):
    R = _ref()
    rng = np.random.RandomState(0)
    for size in (3, 6, 8, 11):
        ref = R.ref_sg_create(size)
        g = oracle.XoSimpleGame()
        for ep in range(20):
            R.ref_sg_reset(ref)
            oracle_lib.xo_sg_reset(C.byref(g), size)
            for _ in range(size + 3):  # keeps acting after the game is over, as the reference allows
                a = int(rng.randint(0, 2))
                assert R.ref_sg_take_action(ref, a) == oracle_lib.xo_sg_act(C.byref(g), a)
                scr = np.zeros(size, np.uint8)
                R.ref_sg_screen(ref, scr.ctypes.data, size)
                assert list(scr) == list(g.state)[:size]
                assert (R.ref_sg_game_over(ref) != 0) == (oracle_lib.xo_sg_game_over(C.byref(g)) != 0)


def test_simple_race_port_vs_compiled_reference(oracle_lib):
    """The C port must equal the reference binary bit for bit (tolerance in BASELINE is 1e-6)."""
    R = _ref()
    for tt, full, hard in [(0, 0, 0), (0, 1, 1), (1, 0, 0), (1, 1, 1)]:
        cfg = _abi.default_config(game=2, track_type=tt, race_full_manouver=full, difficulty=hard)
        ref = R.ref_race_create(tt, 20.0, 100.0, 30.0, full, hard, 1.0, 0)
        o = oracle.XoRace()
        rng = np.random.RandomState(tt * 7 + full)
        for ep in range(40):
            R.ref_race_reset(ref)
            oracle_lib.xo_race_reset(C.byref(cfg), C.byref(o))
            for s in range(300):
                a = int(rng.randint(0, 9 if full else 2))
                st1, ov1 = (C.c_float * 4)(), C.c_int()
                st2, ov2 = (C.c_float * 4)(), C.c_int32()
                r1 = R.ref_race_step(ref, a, st1, C.byref(ov1))
                r2 = oracle_lib.xo_race_act(C.byref(cfg), C.byref(o), a, st2, C.byref(ov2))
                a1 = np.array([r1] + list(st1), np.float32)
                a2 = np.array([r2] + list(st2), np.float32)
                assert (a1.view(np.uint32) == a2.view(np.uint32)).all() and ov1.value == ov2.value, (tt, full, hard, ep, s)
                if ov1.value:
                    break


def test_simple_race_act_rep_vs_compiled_reference(oracle_lib):
    """--act_rep > 1 and --max_steps: the reference's own GameSimulator::take_actions (simulator.cpp:98-108) around the
    compiled SimpleRaceGame against the port's xo_race_take_actions, bit for bit (reward sum, state, game_over code)."""
    R = _ref()
    R.ref_race_take_actions.restype = C.c_float
    R.ref_race_take_actions.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    for tt, full, hard, rep, ms in [(0, 0, 0, 3, 0), (0, 1, 1, 2, 25), (1, 1, 0, 4, 0), (1, 1, 1, 5, 12)]:
        cfg = _abi.default_config(game=2, track_type=tt, race_full_manouver=full, difficulty=hard, max_steps=ms)
        ref = R.ref_race_create(tt, 20.0, 100.0, 30.0, full, hard, 1.0, ms)
        o = oracle.XoRace()
        rng = np.random.RandomState(tt * 7 + full + rep)
        for ep in range(30):
            R.ref_race_reset(ref)
            oracle_lib.xo_race_reset(C.byref(cfg), C.byref(o))
            for s in range(120):
                a = int(rng.randint(0, 9 if full else 2))
                st1, ov1 = (C.c_float * 4)(), C.c_int()
                st2, ov2 = (C.c_float * 4)(), C.c_int32()
                r1 = R.ref_race_take_actions(ref, a, rep, st1, C.byref(ov1))
                r2 = oracle_lib.xo_race_take_actions(C.byref(cfg), C.byref(o), a, rep, st2, C.byref(ov2))
                a1 = np.array([r1] + list(st1), np.float32)
                a2 = np.array([r2] + list(st2), np.float32)
                assert (a1.view(np.uint32) == a2.view(np.uint32)).all() and ov1.value == ov2.value, (tt, full, hard, rep, ep, s)
                if ov1.value:
                    break
        R.ref_race_destroy(ref)


def test_simple_race_random_start_vs_compiled_reference(oracle_lib):
    """--random (simple_race_simulator.cpp:24,267-284): track draw, start position and heading from the thread's
    std::default_random_engine through util::get_rand_range_val -- the compiled reference, its engine re-seeded, against the
    port (libstdc++'s uniform_real_distribution<float> restated), bit for bit over the trajectories that follow each reset."""
    R = _ref()
    R.ref_seed_thread_engine.argtypes = [C.c_uint]
    R.ref_race_set_random(1)
    try:
        for tt, full, hard in [(0, 0, 0), (0, 1, 1), (1, 1, 0), (1, 0, 1)]:
            cfg = _abi.default_config(game=2, track_type=tt, race_full_manouver=full, difficulty=hard, race_random=1)
            ref = R.ref_race_create(tt, 20.0, 100.0, 30.0, full, hard, 1.0, 0)
            rng = np.random.RandomState(31 + tt + 2 * full)
            for seed in [1, 2, 12345, 2147483646, 987654321, 16807, 77]:
                R.ref_seed_thread_engine(seed)
                o = oracle.XoRace()
                o.minstd = seed % 2147483647 or 1   # linear_congruential_engine::seed
                for ep in range(12):   # consecutive episodes: the engine state carries over
                    R.ref_race_reset(ref)
                    oracle_lib.xo_race_reset(C.byref(cfg), C.byref(o))
                    for s in range(40):
                        a = int(rng.randint(0, 9 if full else 2))
                        st1, ov1 = (C.c_float * 4)(), C.c_int()
                        st2, ov2 = (C.c_float * 4)(), C.c_int32()
                        r1 = R.ref_race_step(ref, a, st1, C.byref(ov1))
                        r2 = oracle_lib.xo_race_act(C.byref(cfg), C.byref(o), a, st2, C.byref(ov2))
                        a1 = np.array([r1] + list(st1), np.float32)
                        a2 = np.array([r2] + list(st2), np.float32)
                        assert (a1.view(np.uint32) == a2.view(np.uint32)).all() and ov1.value == ov2.value, (tt, seed, ep, s, a1, a2)
                        if ov1.value:
                            break
            R.ref_race_destroy(ref)
    finally:
        R.ref_race_set_random(0)


def test_step_rules_vs_compiled_xmap(oracle_lib, synthetic_catalog):
    """XAgent::act + XMap::move_item compiled from the reference vs the oracle's move rules, on
    generated maps with random action streams (position, success flag, contacted item)."""
    R = _ref()
    cfg = _abi.default_config(height=7, width=7, n_goals=4, n_blocks=12, rules=0, seed=77, simulator_seed=4)
    o = oracle.Oracle(cfg, synthetic_catalog, 16)
    o.reset()
    rng = np.random.RandomState(3)
    for e in o.envs:
        grid = np.array(list(e.grid)[:49])
        cells = np.nonzero(grid)[0]
        types = np.array([0 if grid[c] == 1 else 2 if grid[c] == 2 else 1 for c in cells], np.int32)
        xs = (cells % 7).astype(np.int32)
        ys = (cells // 7).astype(np.int32)
        m = R.ref_map_create(7, 7, len(cells), types.ctypes.data, xs.ctypes.data, ys.ctypes.data, 1.5707963, 0)
        for s in range(60):
            a = int(rng.randint(0, 4))
            ax, ay, nc, ct = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            yaw = C.c_double()
            ok = R.ref_map_act(m, a, C.byref(ax), C.byref(ay), C.byref(yaw), C.byref(ct), C.byref(nc))
            before = np.array(list(e.grid)[:49])
            tx = e.agent_x + [0, 0, -1, 1][a]
            ty = e.agent_y + [-1, 1, 0, 0][a]
            want_contact = before[ty * 7 + tx] if (0 <= tx < 7 and 0 <= ty < 7) else 0
            r, ov = C.c_float(), C.c_int32()
            oracle_lib.xo_step(C.byref(cfg), C.byref(o.cat_c), C.byref(e), a, 1, C.byref(r), C.byref(ov))
            assert (e.agent_x, e.agent_y, e.action_success) == (ax.value, ay.value, ok)
            assert (nc.value > 0) == (want_contact != 0)
            if nc.value:
                assert grid[cells[ct.value]] == want_contact
        R.ref_map_destroy(m)


def test_rec_group_task_sampling_vs_compiled_reference(oracle_lib):
    """get_extra_info's "task:" field under walls.json needs which XWorldRec task the teacher sampled
    (xworld_b200/csrc/xw_teacher_names.hpp rec_task_of_draw): util::simple_importance_sampling compiled from the reference
    (simulator_util.cpp:56-86) against the restated libstdc++ arithmetic, draw by draw."""
    import parity
    R = _ref()
    H = parity.HostSim.lib()
    H.hs_rec_task_of_draw.argtypes = [C.c_uint32]
    # which thread number will the reference's next fresh thread get?  (its counter is process-wide, simulator_util.cpp:36-47)
    out = (C.c_int * 1)()
    R.ref_rand_ind_threads(5, 1, 1000000, out)
    cur = None
    for t in range(1, 4096):
        st = C.c_uint32(oracle_lib.xo_minstd_seed_for_thread(5, t))
        if oracle_lib.xo_get_rand_ind(C.byref(st), 1000000) == out[0]:
            cur = t
            break
    assert cur is not None
    acc = (C.c_double * 12)(1, 2, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14)
    n = 20000
    idx = (C.c_int * n)()
    R.ref_importance_sampling.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    R.ref_importance_sampling(5, acc, 12, n, idx)
    x = oracle_lib.xo_minstd_seed_for_thread(5, cur + 1)
    seen = set()
    for i in range(n):
        x = (x * 16807) % 2147483647
        assert H.hs_rec_task_of_draw(x) == idx[i], i
        seen.add(idx[i])
    assert seen == set(range(12))
