"""The parity tests proper: the CUDA engine, called through the C ABI, against the oracle on the
same seeded inputs -- bit-exact state / reward bits / game_over, pixel-exact frames."""
import ctypes as C

import numpy as np
import pytest

import oracle
import parity
from test_oracle_render import golden_cases
from xworld_b200 import Simulator, _abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend_cls():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from gpu_backend import EngineBackend
    return EngineBackend


@pytest.mark.parametrize("name", sorted(parity.CONFIGS))
def test_engine_matches_oracle(name, backend_cls, synthetic_catalog):
    cfg = parity.make_cfg(name)
    n = 1024 if cfg.height <= 8 else 512
    eng = backend_cls(cfg, synthetic_catalog, n)
    orc = oracle.Oracle(cfg, synthetic_catalog, n, threads=8)
    stats = parity.run_parity(eng, orc, 128, render_every=16, check_state_every=8)
    assert stats["frames"] >= 8 * n
    assert eng.sim.launch_count() > 128


def test_config2_full_size_256_steps(backend_cls, synthetic_catalog):
    """BASELINE config 2: 7x7, 84x84x3, 4096 envs, bit-exact over 256 steps."""
    cfg = parity.make_cfg("c2_nav3d_7x7_84")
    eng = backend_cls(cfg, synthetic_catalog, 4096)
    orc = oracle.Oracle(cfg, synthetic_catalog, 4096, threads=8)
    stats = parity.run_parity(eng, orc, 256, render_every=64, check_state_every=32)
    assert stats["events"].get(_abi.XW_SUCCESS, 0) > 0 and stats["events"].get(_abi.XW_DEAD, 0) > 0


def test_auto_reset_and_act_rep(backend_cls, synthetic_catalog):
    cfg = parity.make_cfg("c2_nav3d_7x7_84", auto_reset=1, max_steps=30)
    eng = backend_cls(cfg, synthetic_catalog, 2048)
    orc = oracle.Oracle(cfg, synthetic_catalog, 2048, threads=8)
    st = parity.run_parity(eng, orc, 100, render_every=25, auto_reset=True, check_state_every=10)
    assert st["events"].get(_abi.XW_MAX_STEP, 0) > 0
    cfg = parity.make_cfg("c3_nav2d_11x11_84", auto_reset=1)
    eng = backend_cls(cfg, synthetic_catalog, 512)
    orc = oracle.Oracle(cfg, synthetic_catalog, 512, threads=8)
    parity.run_parity(eng, orc, 300, render_every=100, auto_reset=True, act_rep=2, check_state_every=50)


@pytest.mark.parametrize("kw,n,steps", [
    (dict(curriculum=0.1, curriculum_check_period=4, auto_reset=1, max_steps=40), 2048, 400),  # warp-per-env k_reset
    (dict(curriculum=0.02, start_level=2, curriculum_check_period=3), 1024, 300),              # offset-1 levels
    (dict(curriculum=0.02, start_level=4, curriculum_check_period=2, auto_reset=1), 1024, 300),
    (dict(curriculum=0.9, max_steps_factor=1, auto_reset=1), 32, 10000),                        # windows fill and roll
])
def test_curriculum_levels(kw, n, steps, backend_cls, synthetic_catalog):
    """SURVEY 8f-3 (XWorldNav.py:36-58): per-env levels, padded worlds, result windows, the displaced-referent
    rule of the padded levels -- state, rewards, game_over and frames bit-exact against the oracle."""
    cfg = parity.make_cfg("curriculum_nav3d_8x8_96", **kw)
    eng = backend_cls(cfg, synthetic_catalog, n)
    orc = oracle.Oracle(cfg, synthetic_catalog, n, threads=8)
    parity.run_parity(eng, orc, steps, render_every=steps // 4, check_state_every=steps // 8,
                      auto_reset=bool(kw.get("auto_reset")))
    lv = eng.field("level")
    assert lv.min() >= kw.get("start_level", 0)
    if kw.get("max_steps_factor") == 1:
        wl = eng.field("win_len")
        assert wl.max() == 200 and wl.min() >= 150
    elif "start_level" not in kw:
        assert lv.max() >= 1
    if kw.get("start_level") == 4:
        assert lv.max() == 5


def test_golden_frames_from_real_opencv(backend_cls):
    """The GPU compositor vs frames the real OpenCV produced from the reference call sequence."""
    n = 0
    for tag, cfg, cat, grid, gi, want in golden_cases():
        eng = backend_cls(cfg, cat, 1)
        eng.reset()
        eng.sim.set_field("grid", np.asarray(grid, np.uint8)[None, :])
        icons = np.zeros((1, _abi.XW_MAX_GOALS), np.int32)
        icons[0, :4] = gi
        eng.sim.set_field("goal_icon", icons)
        got = eng.render()[0]
        assert (got == want).all(), (tag, int((got != want).sum()))
        n += 1
    assert n == 12


def test_generic_render_path_matches(backend_cls, synthetic_catalog):
    """Frame sizes the shared-memory compositor does not take (width % 4 != 0) use the generic kernel."""
    cfg = parity.make_cfg("c2_nav3d_7x7_84", out_h=50, out_w=70)
    eng = backend_cls(cfg, synthetic_catalog, 64)
    orc = oracle.Oracle(cfg, synthetic_catalog, 64, threads=4)
    parity.run_parity(eng, orc, 20, render_every=5)


def test_full_size_properties_c3(backend_cls, synthetic_catalog):
    """BASELINE config 3 at full size (65536 envs): size-independent properties -- every frame is a
    pure function of (grid, goal icons): envs sampled at random are pixel-exact vs the oracle; the
    white fraction equals the empty-cell fraction; re-rendering is idempotent."""
    cfg = parity.make_cfg("c3_nav2d_11x11_84", auto_reset=1)
    n = 65536
    eng = backend_cls(cfg, synthetic_catalog, n)
    eng.reset()
    for s in range(20):
        eng.step(parity.actions_for(s, n, 4), render=False)
    f1 = eng.render()
    f2 = eng.render()
    assert (f1 == f2).all()
    grid, icons = eng.field("grid"), eng.field("goal_icon")
    assert ((grid == _abi.XW_CELL_AGENT).sum(axis=1) == 1).all() and ((grid == _abi.XW_CELL_BLOCK).sum(axis=1) == 30).all()
    rng = np.random.RandomState(0)
    lib = oracle.lib()
    for i in rng.choice(n, 64, replace=False):
        e = oracle.XoEnv()
        lib.xo_env_init(C.byref(cfg), C.byref(e), 0)
        for c, v in enumerate(grid[i]):
            e.grid[c] = int(v)
        for k in range(4):
            e.goal_icon[k] = int(icons[i, k])
        want = np.zeros((3, 84, 84), np.uint8)
        lib.xo_render(C.byref(cfg), C.byref(synthetic_catalog.as_c()), C.byref(e), want.ctypes.data)
        assert (f1[i] == want).all(), i


def test_sharded_engine_equals_single_batch(backend_cls, synthetic_catalog):
    from xworld_b200.sharding import shard_range
    n = 512
    full = backend_cls(parity.make_cfg("c2_nav3d_7x7_84"), synthetic_catalog, n)
    full.reset()
    lo, hi = shard_range(n, 1, 2)
    part = backend_cls(parity.make_cfg("c2_nav3d_7x7_84", env_id_offset=lo), synthetic_catalog, hi - lo)
    part.reset()
    for s in range(20):
        a = parity.actions_for(s, n, 4)
        r1, o1, f1 = full.step(a, render=(s == 19))
        r2, o2, f2 = part.step(a[lo:hi], render=(s == 19))
        assert (r1[lo:hi].view(np.uint32) == r2.view(np.uint32)).all() and (o1[lo:hi] == o2).all()
    assert (f1[lo:hi] == f2).all()


def test_python_api_single_env(synthetic_catalog, tmp_path):
    """The reference's Python surface with n_envs == 1 (py_simulator.cpp:310-329)."""
    conf = tmp_path / "navigation2d.json"
    conf.write_text('{"item_path": "images", "map": "XWorldNav", "task_groups": {"XWorld3DNav": {"weight": 1, '
                    '"schedule": "random", "tasks": {"XWorld3DNavTarget": 1, "XWorld3DNavTargetNear": 1, '
                    '"XWorld3DNavTargetBetween": 1, "XWorld3DNavTargetDirection": 1, "XWorld3DNavTargetAvoid": 1}}}}')
    opts = {"xwd_conf_path": str(conf), "task_mode": "lang_acquisition", "color": True, "catalog": synthetic_catalog,
            "simulator_seed": 1, "seed": 1234}
    sim = Simulator.create("xworld", opts)
    assert sim.get_num_actions() == 4 and sim.get_screen_out_dimensions() == [96, 96, 3, 1]
    sim.reset_game()
    assert sim.game_over() == "alive" and sim.get_lives() == 1
    st = sim.get_state()
    assert len(st["screen"]) == 3 * 96 * 96 and 0.0 <= min(st["screen"]) and max(st["screen"]) <= 1.0
    total = 0.0
    for s in range(30):
        r = sim.take_actions({"action": int(s % 4)}, 1, False)
        assert isinstance(r, float)
        total += r
        if sim.game_over() != "alive":
            break
    assert sim.get_num_steps() == s + 1
    # the reference's own defaults (py_simulator.cpp:124-136): task_mode one_channel, color false -> one gray plane
    dflt = Simulator.create("xworld", {"xwd_conf_path": str(conf), "catalog": synthetic_catalog})
    assert dflt.cfg.task_mode == _abi.XW_TASK_ONE_CHANNEL and dflt.get_screen_out_dimensions() == [96, 96, 1, 1]
    dflt.reset_game()
    assert len(dflt.get_state()["screen"]) == 96 * 96
    for s2 in range(12):
        dflt.take_actions({"action": s2 % 4, "pred_sentence": ""}, 1, False)
        assert dflt.game_over() == "alive"   # one_channel: only --max_steps ends a session (xworld_simulator.cpp:192-193)
    # invalid action: flagged, never aborts (the reference CHECK-fails, xworld_simulator.cpp:254)
    sim.take_actions({"action": 7})
    assert sim.get_field("error")[0] == -4


def test_context_frames(backend_cls, synthetic_catalog):
    """--context K: K frames per env, oldest first, newest last (simulator.cpp:51-85)."""
    cfg = parity.make_cfg("c2_nav3d_7x7_84", context=3)
    n = 128
    eng = backend_cls(cfg, synthetic_catalog, n)
    one = backend_cls(parity.make_cfg("c2_nav3d_7x7_84"), synthetic_catalog, n)
    eng.reset()
    one.reset()
    hist = [one.render().copy()]
    scr = eng.sim.screen().cpu().numpy()
    assert (scr[:, :6] == 0).all() and (scr[:, 6:] == hist[0]).all()
    for s in range(4):
        a = parity.actions_for(s, n, 4)
        eng.step(a, render=True)
        _, _, f = one.step(a, render=True)
        hist.append(f.copy())
    scr = eng.sim.screen().cpu().numpy()
    assert (scr[:, 0:3] == hist[-3]).all() and (scr[:, 3:6] == hist[-2]).all() and (scr[:, 6:9] == hist[-1]).all()


def test_simple_race_vs_oracle_1e6(backend_cls):
    """BASELINE config 5 (tol 1e-6 on reward and the 4-float state); reports exact-bit agreement."""
    import torch
    lib = oracle.lib()
    # (the third: --act_rep 3, simulator.cpp:98-108; the last two: --random start states, simple_race_simulator.cpp:267-284)
    for tt, full, hard, rep, rnd in [(0, 0, 0, 1, 0), (1, 1, 1, 1, 0), (1, 1, 0, 3, 0), (0, 1, 0, 1, 1), (1, 1, 1, 2, 1)]:
        cfg = _abi.default_config(game=_abi.XW_GAME_SIMPLE_RACE, track_type=tt, race_full_manouver=full, difficulty=hard,
                                  auto_reset=1, race_random=rnd, simulator_seed=9)
        n = 4096
        eng = backend_cls(cfg, None, n)
        eng.reset()
        orcs = (oracle.XoRace * n)()
        for i, o in enumerate(orcs):
            o.minstd = lib.xo_minstd_seed_for_thread(9, i + 1)
            lib.xo_race_reset(C.byref(cfg), C.byref(o))
        rng = np.random.RandomState(11)
        exact = total = 0
        for s in range(60):
            a = rng.randint(0, 9 if full else 2, n).astype(np.int32)
            r, ov, _ = eng.step(a, act_rep=rep)
            st = eng.field("state")
            r2 = np.zeros(n, np.float32)
            st2 = np.zeros((n, 4), np.float32)
            ov2 = np.zeros(n, np.int32)
            for i in range(n):
                buf, o2 = (C.c_float * 4)(), C.c_int32()
                r2[i] = lib.xo_race_take_actions(C.byref(cfg), C.byref(orcs[i]), int(a[i]), rep, buf, C.byref(o2))
                st2[i] = list(buf)
                ov2[i] = o2.value
                if o2.value:
                    lib.xo_race_reset(C.byref(cfg), C.byref(orcs[i]))
            assert np.abs(r - r2).max() <= 1e-6 and np.abs(st - st2).max() <= 1e-6 and (ov == ov2).all(), (tt, s)
            exact += int((r.view(np.uint32) == r2.view(np.uint32)).sum())
            total += n
        assert exact >= 0.999 * total, (exact, total)


def test_step_seq_equals_single_steps(backend_cls, synthetic_catalog):
    """xw_step_seq (K take_actions calls per launch, include/xworld_b200.h) against K xw_step calls on a twin handle: reward
    bits, game_over and the final state, for SimpleRace (one launch, car in registers; --random and act_rep included) and xworld."""
    import torch
    K = 24
    for kw, n, n_act, rep in [(dict(game=_abi.XW_GAME_SIMPLE_RACE, track_type=1, race_full_manouver=1, difficulty=1, auto_reset=1), 5000, 9, 1),
                              (dict(game=_abi.XW_GAME_SIMPLE_RACE, track_type=0, race_full_manouver=1, race_random=1, simulator_seed=3,
                                    auto_reset=1, max_steps=17), 3000, 9, 2),
                              (dict(height=7, width=7, n_goals=4, n_blocks=12, rules=0, auto_reset=1, seed=5, simulator_seed=2), 2048, 4, 1)]:
        cfg = _abi.default_config(**kw)
        cat = None if cfg.game == _abi.XW_GAME_SIMPLE_RACE else synthetic_catalog
        a_eng, b_eng = backend_cls(cfg, cat, n), backend_cls(cfg, cat, n)
        a_eng.reset(); b_eng.reset()
        rng = np.random.RandomState(2)
        acts = rng.randint(0, n_act, (K, n)).astype(np.int32)
        r1 = np.zeros((K, n), np.float32); o1 = np.zeros((K, n), np.int32)
        for k in range(K):
            r1[k], o1[k], _ = a_eng.step(acts[k], act_rep=rep)
        sim = b_eng.sim
        with torch.cuda.device(sim._dev):
            da = torch.from_numpy(acts).cuda()
            dr = torch.zeros((K, n), dtype=torch.float32, device="cuda")
            do = torch.zeros((K, n), dtype=torch.int32, device="cuda")
            rc = sim._lib.xw_step_seq(sim._h, da.data_ptr(), K, rep, dr.data_ptr(), do.data_ptr(), sim._stream())
            assert rc == 0, sim._lib.xw_last_error()
            r2, o2 = dr.cpu().numpy(), do.cpu().numpy()
        assert (r1.view(np.uint32) == r2.view(np.uint32)).all() and (o1 == o2).all()
        for f in (["state"] if cfg.game == _abi.XW_GAME_SIMPLE_RACE else ["grid", "agent_x", "agent_y", "num_steps", "episode"]):
            assert (a_eng.field(f).view(np.uint8) == b_eng.field(f).view(np.uint8)).all(), f


def test_action_none_and_batch_client(backend_cls, synthetic_catalog):
    """XW_ACTION_NONE through the C ABI (envs that sit a step out are untouched, their output slots keep their last
    values), then the same engine behind the reference's TCP protocol: wire.BatchClient with one connection per env."""
    import socket
    import struct
    import threading
    from xworld_b200 import wire
    for name in ("c2_nav3d_7x7_84", "c3_nav2d_11x11_84"):
        cfg = parity.make_cfg(name)
        eng = backend_cls(cfg, synthetic_catalog, 512)
        orc = oracle.Oracle(cfg, synthetic_catalog, 512, threads=8)
        assert parity.run_partial_parity(eng, orc, 60, render_every=20) > 10000
    # ---- the wire protocol in front of a GPU batch
    n = 8
    cfg = parity.make_cfg("c2_nav3d_7x7_84")
    eng = backend_cls(cfg, synthetic_catalog, n)
    orc = oracle.Oracle(cfg, synthetic_catalog, n, threads=1)
    lsocks = []
    for _ in range(n):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        s.listen(1)
        lsocks.append(s)
    box = {}

    def run():
        box["c"] = wire.BatchClient(eng.sim, [s.getsockname()[1] for s in lsocks])
        box["c"].serve()

    th = threading.Thread(target=run, daemon=True)
    th.start()
    conns = [s.accept()[0] for s in lsocks]

    def read_msg(c):
        b = b""
        while len(b) < 8:
            b += c.recv(8 - len(b))
        size = struct.unpack("<Q", b)[0]
        body = b""
        while len(body) < size:
            body += c.recv(size - len(body))
        return body

    for c in conns:
        c.sendall(wire.compose_request("reset"))
    for c in conns:
        body = read_msg(c)
        assert body[8:13] == b"reset" and struct.unpack("<iii", body[14:26]) == (4, 0, 1)
        assert struct.unpack("<QQQddd", body[26:74]) == (84, 84, 3, 7.0, 7.0, 0.0)
    orc.reset()
    rng = np.random.RandomState(1)
    for it in range(40):
        who = [i for i in range(n) if rng.rand() < 0.6] or [3]
        acts = {i: int(rng.randint(0, 4)) for i in who}
        for i in who:
            conns[i].sendall(wire.compose_request("take_actions", {"action": acts[i], "pred_sentence": ""}))
        for i in who:
            body = read_msg(conns[i])
            assert body[8:20] == b"take_actions"
            r, num_steps, over, lives, ok = struct.unpack("<fqiiB", body[21:42])
            r2, o2 = C.c_float(), C.c_int32()
            assert orc.L.xo_step(C.byref(orc.cfg), C.byref(orc.cat_c), C.byref(orc.envs[i]), acts[i], 1, C.byref(r2), C.byref(o2)) == 0
            e = orc.envs[i]
            assert (np.float32(r), num_steps, over, lives, ok) == (np.float32(r2.value), e.num_steps, o2.value, 0 if o2.value else 1,
                                                                   e.action_success), (it, i)
            if over:
                conns[i].sendall(wire.compose_request("reset"))
                read_msg(conns[i])
                m = np.zeros(n, np.uint8)
                m[i] = 1
                orc.reset(m)
    conns[2].sendall(wire.compose_request("get_state", reward=0.0))
    st = wire.decode_packet(read_msg(conns[2])[8 + 10:])
    assert (st["screen"].reshape(3, 84, 84) == orc.render([2])[0]).all() and isinstance(st["sentence"], str)
    conns[5].sendall(wire.compose_request("get_extra_info"))
    assert b"height:7,width:7" in read_msg(conns[5])
    for c in conns:
        c.sendall(wire.compose_request("stop"))
    th.join(timeout=20)
    assert not th.is_alive() and box["c"].batches < box["c"].steps_served
    parity.compare_state(eng, orc, "after the wire session")


def test_checkpoint_resume_is_bit_exact(backend_cls, synthetic_catalog):
    """Simulator.state_dict / load_state_dict: a fresh handle loaded with a snapshot continues exactly like the
    original (state, rewards, game_over, frames), including the curriculum windows and auto-reset episodes."""
    for name, kw in (("curriculum_nav3d_8x8_96", dict(auto_reset=1, curriculum_check_period=3)), ("c3_nav2d_11x11_84", dict(auto_reset=1))):
        cfg = parity.make_cfg(name, **kw)
        a = backend_cls(cfg, synthetic_catalog, 512)
        a.reset()
        for s in range(150):
            a.step(parity.actions_for(s, 512, 4))
        snap = a.sim.state_dict()
        b = backend_cls(cfg, synthetic_catalog, 512)
        b.sim.load_state_dict(snap)
        for s in range(150, 260):
            act = parity.actions_for(s, 512, 4)
            r1, o1, f1 = a.step(act, render=(s % 50 == 0))
            r2, o2, f2 = b.step(act, render=(s % 50 == 0))
            assert (r1.view(np.uint32) == r2.view(np.uint32)).all() and (o1 == o2).all(), (name, s)
            if f1 is not None:
                assert (f1 == f2).all(), (name, s)
        s1, s2 = a.sim.state_dict(), b.sim.state_dict()
        for k in s1:
            assert (np.asarray(s1[k]) == np.asarray(s2[k])).all(), (name, k)


def test_simple_race_full_size_c5(backend_cls):
    """BASELINE config 5 at its full size: 1,048,576 envs.  Envs start in the same state (random=false), so envs fed
    the same action stream must stay bit-identical; 64 streams, each checked against the oracle (tol 1e-6)."""
    lib = oracle.lib()
    n, groups, steps = 1 << 20, 64, 120
    for tt, full in [(0, 0), (1, 1)]:
        cfg = _abi.default_config(game=_abi.XW_GAME_SIMPLE_RACE, track_type=tt, race_full_manouver=full, auto_reset=1)
        eng = backend_cls(cfg, None, n)
        eng.reset()
        orcs = (oracle.XoRace * groups)()
        for o in orcs:
            lib.xo_race_reset(C.byref(cfg), C.byref(o))
        rng = np.random.RandomState(3 + tt)
        gid = np.arange(n) % groups
        n_over = 0
        for s in range(steps):
            ga = rng.randint(0, 9 if full else 2, groups).astype(np.int32)
            r, ov, _ = eng.step(ga[gid])
            st = eng.field("state")
            r2, st2, ov2 = np.zeros(groups, np.float32), np.zeros((groups, 4), np.float32), np.zeros(groups, np.int32)
            for g in range(groups):
                buf, o2 = (C.c_float * 4)(), C.c_int32()
                r2[g] = lib.xo_race_act(C.byref(cfg), C.byref(orcs[g]), int(ga[g]), buf, C.byref(o2))
                st2[g], ov2[g] = list(buf), o2.value
                if o2.value:
                    lib.xo_race_reset(C.byref(cfg), C.byref(orcs[g]))
            # every env equals the first env of its group, bit for bit
            assert (r.view(np.uint32) == r.view(np.uint32)[:groups][gid]).all(), s
            assert (st.view(np.uint32) == st.view(np.uint32)[:groups][gid]).all() and (ov == ov[:groups][gid]).all(), s
            assert np.abs(r[:groups] - r2).max() <= 1e-6 and np.abs(st[:groups] - st2).max() <= 1e-6 and (ov[:groups] == ov2).all(), s
            n_over += int((ov2 != 0).sum())
        assert n_over > 0 or tt == 0


@pytest.mark.parametrize("name,n", [("c3_nav2d_11x11_84", 65536), ("c2_nav3d_7x7_84", 65536), ("c4_nav3d_15x15_128", 32768),
                                    ("ref_nav3d_8x8_96", 16384)])
def test_painter_equals_plan_compositor_full_size(name, n, synthetic_catalog, monkeypatch):
    """Two independent render kernels -- the sparse painter (default) and the dense plan compositor
    (XW_RENDER_MODE=sb) -- on the same states at the BASELINE env counts: every byte of every frame equal."""
    import torch
    assert torch.cuda.is_available()
    from gpu_backend import EngineBackend
    cfg = parity.make_cfg(name, auto_reset=1)
    engines = {}
    for mode in ("sp", "sb"):
        monkeypatch.setenv("XW_RENDER_MODE", mode)
        engines[mode] = EngineBackend(cfg, synthetic_catalog, n)
        engines[mode].reset()
    monkeypatch.delenv("XW_RENDER_MODE")
    assert engines["sp"].sim.render_kernel() == 3 and engines["sb"].sim.render_kernel() == 1
    n_act = engines["sp"].sim.get_num_actions()
    for s in range(12):
        a = parity.actions_for(s, n, n_act)
        for e in engines.values():
            e.step(a, render=False)
    for e in engines.values():
        sim = e.sim
        with torch.cuda.device(sim._dev):
            assert sim._lib.xw_render(sim._h, sim._screen.data_ptr(), sim._stream()) == 0
    torch.cuda.synchronize()
    a, b = engines["sp"].sim._screen, engines["sb"].sim._screen
    assert a.shape == b.shape and bool(torch.equal(a, b)), int((a != b).sum())
    white = float((a == 255).float().mean())
    assert 0.3 < white < 0.95  # a maze: mostly white, never blank


@pytest.mark.parametrize("side,out,n", [(16, 128, 37), (16, 96, 300), (3, 24, 5), (5, 40, 149), (12, 96, 1185), (13, 104, 64)])
def test_painter_odd_shapes_vs_oracle(side, out, n, backend_cls, synthetic_catalog):
    """Map sides and env counts off the beaten path: 16x16 maps (two grid words per lane in the painter's cell
    loader), fewer envs than warp groups, env counts that are not a multiple of the group count."""
    blocks = {16: 60, 3: 1, 5: 4, 12: 36, 13: 42}[side]
    cfg = _abi.default_config(height=side, width=side, n_goals=min(4, side - 1), n_blocks=blocks, rules=_abi.XW_RULES_NAV3D,
                              out_h=out, out_w=out, seed=77, simulator_seed=3)
    eng = backend_cls(cfg, synthetic_catalog, n)
    orc = oracle.Oracle(cfg, synthetic_catalog, n, threads=8)
    stats = parity.run_parity(eng, orc, 24, render_every=6, check_state_every=6)
    assert stats["frames"] >= 4 * n


def test_step_hd_host_buffers_equal_device_path(synthetic_catalog):
    """xw_step_hd (host actions in, host reward / game_over out, read-back overlapped with the render kernel)
    against xw_step on a twin engine: page-locked buffers (used in place) and pageable ones (staged)."""
    import torch
    from gpu_backend import EngineBackend
    cfg = parity.make_cfg("c3_nav2d_11x11_84", auto_reset=1)
    n = 4096
    a_eng, b_eng = EngineBackend(cfg, synthetic_catalog, n), EngineBackend(cfg, synthetic_catalog, n)
    a_eng.reset(); b_eng.reset()
    sim = b_eng.sim
    for pinned in (True, False):
        h_act = torch.zeros(n, dtype=torch.int32)
        h_rew, h_over = torch.zeros(n, dtype=torch.float32), torch.zeros(n, dtype=torch.int32)
        if pinned:
            h_act, h_rew, h_over = h_act.pin_memory(), h_rew.pin_memory(), h_over.pin_memory()
        for s in range(30):
            a = parity.actions_for(s + (100 if pinned else 0), n, 4)
            r1, o1, f1 = a_eng.step(a, render=True)
            h_act.copy_(torch.from_numpy(a))
            with torch.cuda.device(sim._dev):
                rc = sim._lib.xw_step_hd(sim._h, h_act.data_ptr(), 1, h_rew.data_ptr(), h_over.data_ptr(), sim._screen.data_ptr())
            assert rc == 0, sim._lib.xw_last_error()
            assert (h_rew.numpy().view(np.uint32) == r1.view(np.uint32)).all() and (h_over.numpy() == o1).all()
            torch.cuda.synchronize()
            assert (sim._screen.cpu().numpy() == f1).all()
    # the pipelined form: reward / game_over on return, frames after xw_wait_frames / xw_sync; two frame buffers
    h_act, h_rew, h_over = (torch.zeros(n, dtype=torch.int32).pin_memory(), torch.zeros(n, dtype=torch.float32).pin_memory(),
                            torch.zeros(n, dtype=torch.int32).pin_memory())
    bufs = [torch.empty_like(sim._screen), torch.empty_like(sim._screen)]
    prev = None
    for s in range(40):
        a = parity.actions_for(s + 500, n, 4)
        r1, o1, f1 = a_eng.step(a, render=True)
        h_act.copy_(torch.from_numpy(a))
        with torch.cuda.device(sim._dev):
            rc = sim._lib.xw_step_hd_async(sim._h, h_act.data_ptr(), 1, h_rew.data_ptr(), h_over.data_ptr(), bufs[s & 1].data_ptr())
            assert rc == 0, sim._lib.xw_last_error()
            assert (h_rew.numpy().view(np.uint32) == r1.view(np.uint32)).all() and (h_over.numpy() == o1).all()
            if prev is not None:  # the previous step's frames, read by a consumer stream while this step renders
                assert (bufs[(s - 1) & 1].cpu().numpy() == prev).all()
            assert sim._lib.xw_wait_frames(sim._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
        prev = f1
    assert sim._lib.xw_sync(sim._h) == 0
    assert (bufs[39 & 1].cpu().numpy() == prev).all()


@pytest.mark.parametrize("name", ["c2_nav3d_7x7_84", "c3_nav2d_11x11_84"])
def test_sentences_follow_the_task_state(name, backend_cls, synthetic_catalog):
    """Simulator.sentences(): the teacher's sentence of every env agrees with the env's task state -- the bound goal
    names / direction / colour are in it (navigation2d.json tasks: while the episode runs; walls.json: in the step
    that issued the command), "Well done !" exactly where the step reached the goal."""
    cfg = parity.make_cfg(name, auto_reset=1)
    n = 2048
    eng = backend_cls(cfg, synthetic_catalog, n)
    eng.reset()
    cat, sim = synthetic_catalog, eng.sim
    seen = set()
    for s in range(12):
        if s:
            eng.step(parity.actions_for(s, n, sim.get_num_actions()))
        sent = sim.sentences()
        f = {k: sim.get_field(k) for k in ("task", "stage", "event", "aux0", "aux1", "goal_name", "goal_icon", "steps_in_task")}
        assert len(sent) == n
        for e in range(0, n, 7):
            t, a0 = int(f["task"][e]), int(f["aux0"][e])
            nm = lambda g: cat.names[int(f["goal_name"][e][g])]
            if cfg.rules == _abi.XW_RULES_NAV3D:
                if f["stage"][e] == _abi.XW_STAGE_NAVIGATION:
                    assert sent[e], e
                    if t == 2:
                        assert "between %s and %s" % (nm(a0 & 15), nm(a0 >> 4)) in sent[e]
                    else:
                        assert nm(a0) in sent[e].split() or nm(a0) in sent[e]
                    if t == 3:
                        assert {1: "front", 2: "behind", 3: "left", 4: "right"}[int(f["aux1"][e])] in sent[e]
                    seen.add(t)
            else:
                if f["event"][e] == _abi.XW_EVENT_CORRECT_GOAL:
                    assert sent[e] == "Well done !"
                elif f["stage"][e] == _abi.XW_STAGE_NAVIGATION and f["steps_in_task"][e] == 0:
                    assert nm(a0) in sent[e]
                    if t == 2:
                        assert cat.icon_meta[int(f["goal_icon"][e][a0])]["color"] in sent[e]
                    seen.add(t)
                else:
                    assert sent[e] == ""
    assert seen >= ({0, 1, 2, 3, 4} if cfg.rules == _abi.XW_RULES_NAV3D else {0, 2})
    one = sim.sentences([5])
    assert one == [sim.sentences()[5]]


def _fpv_scene(eng, grid, gi, pose, agent_yaw, W):
    """Write one first-person scene into env 0 of `eng` through xw_set_field."""
    sim = eng.sim
    g = np.asarray(grid, np.uint8)
    sim.set_field("grid", g[None, :])
    c = int(np.nonzero(g == _abi.XW_CELL_AGENT)[0][0])
    sim.set_field("agent_x", [c % W])
    sim.set_field("agent_y", [c // W])
    sim.set_field("facing", [oracle.lib().xo_facing_dir(agent_yaw)])
    pad = lambda v, dt: np.concatenate([np.asarray(v, dt), np.zeros(_abi.XW_MAX_GOALS - len(v), dt)])[None, :]
    sim.set_field("goal_icon", pad(gi, np.int32))
    sim.set_field("goal_yaw", pad([int(round(p[0] / (1.5707963 * 4) * 4096)) for p in pose], np.uint16))
    sim.set_field("goal_scale", pad([p[1] for p in pose], np.float64))
    sim.set_field("goal_offset", pad([p[2] for p in pose], np.float64))


def test_fpv_golden_frames_from_real_opencv(backend_cls):
    """the first-person kernels (k_fpv_warp_goals + k_render_fpv / k_render_fpv_generic) vs frames the real
    OpenCV produced from the reference call sequence, incl. the reference-native 84x84 view of an 11x11 map."""
    from test_oracle_fpv import fpv_golden_cases
    n = 0
    kernels = set()
    for tag, cfg, cat, grid, gi, pose, agent_yaw, want in fpv_golden_cases():
        eng = backend_cls(cfg, cat, 1)
        eng.reset()
        _fpv_scene(eng, grid, gi, pose, agent_yaw, cfg.width)
        got = eng.render()[0]
        assert (got == want).all(), (tag, n, int((got != want).sum()))
        kernels.add(eng.sim.render_kernel())
        n += 1
    assert n == 42 and kernels == {4, 5, 6}   # the any-size kernel, the shared-memory one and the cell-block one all ran


def test_fpv_streaming_goal_kernel_same_frames(backend_cls, synthetic_catalog, monkeypatch):
    """XW_FPV_STREAM=1 (the goal kernel beside the frame kernel: ready bits, launch sequence numbers, TMA ring) paints the
    same frames as the default (goal kernel after the frame kernel), step after step with auto-reset."""
    cfg = _abi.default_config(height=11, width=11, n_goals=4, n_blocks=30, rules=1, visible_radius=7, max_steps=40, auto_reset=1,
                              seed=21, simulator_seed=3)
    n = 3000
    a_eng = backend_cls(cfg, synthetic_catalog, n)
    monkeypatch.setenv("XW_FPV_STREAM", "1")
    b_eng = backend_cls(cfg, synthetic_catalog, n)
    monkeypatch.delenv("XW_FPV_STREAM")
    assert a_eng.sim.render_kernel() == 6 and b_eng.sim.render_kernel() == 6
    a_eng.reset(); b_eng.reset()
    rng = np.random.RandomState(4)
    for s in range(60):
        a = rng.randint(0, 6, n).astype(np.int32)
        r1, o1, f1 = a_eng.step(a, render=True)
        r2, o2, f2 = b_eng.step(a, render=True)
        assert (r1.view(np.uint32) == r2.view(np.uint32)).all() and (o1 == o2).all()
        assert (f1 == f2).all(), (s, int((f1 != f2).sum()))


def test_fpv_auto_reset_and_context(backend_cls, synthetic_catalog):
    """Auto-reset in the first-person view: the goal icons of re-started episodes are warped by the list-mode launch."""
    cfg = parity.make_cfg("fpv_nav3d_8x8_vr3_84", auto_reset=1, max_steps=30)
    eng = backend_cls(cfg, synthetic_catalog, 2048)
    orc = oracle.Oracle(cfg, synthetic_catalog, 2048, threads=8)
    st = parity.run_parity(eng, orc, 120, render_every=30, auto_reset=True, check_state_every=15)
    assert st["events"].get(_abi.XW_MAX_STEP, 0) > 0 and st["events"].get(_abi.XW_SUCCESS, 0) > 0
    cfg = parity.make_cfg("fpv_nav2d_11x11_vr7_84", auto_reset=1, max_steps=40, context=2)
    eng = backend_cls(cfg, synthetic_catalog, 256)
    one = backend_cls(parity.make_cfg("fpv_nav2d_11x11_vr7_84", auto_reset=1, max_steps=40), synthetic_catalog, 256)
    eng.reset()
    one.reset()
    prev = one.render().copy()
    n_new = 0
    for s in range(50):
        a = parity.actions_for(s, 256, 6)
        ep0 = eng.field("episode").copy()
        eng.step(a, render=True)
        _, _, f = one.step(a, render=True)
        scr = eng.sim.screen().cpu().numpy()
        new = eng.field("episode") != ep0   # auto-reset in this step: the context of a new game starts zero-filled
        n_new += int(new.sum())
        assert (scr[~new, 0:3] == prev[~new]).all() and (scr[new, 0:3] == 0).all() and (scr[:, 3:6] == f).all(), s
        prev = f.copy()
    assert n_new >= 256


def test_fpv_full_size_properties(backend_cls, synthetic_catalog):
    """BASELINE config 3's map in its reference-native 84x84 first-person view (11x11, visible_radius 7) at 65,536 envs:
    re-rendering is idempotent; 64 envs sampled at random are pixel-exact against the oracle rendering the same state."""
    cfg = parity.make_cfg("fpv_nav2d_11x11_vr7_84", auto_reset=1)
    n = 65536
    eng = backend_cls(cfg, synthetic_catalog, n)
    eng.reset()
    for s in range(30):
        eng.step(parity.actions_for(s, n, 6), render=False)
    f1 = eng.render()
    f2 = eng.render()
    assert (f1 == f2).all()
    F = {k: eng.field(k) for k in ("grid", "goal_icon", "goal_yaw", "goal_scale", "goal_offset", "agent_x", "agent_y", "facing")}
    assert set(np.unique(F["facing"])) == {0, 1, 2, 3}
    lib = oracle.lib()
    rng = np.random.RandomState(0)
    yaw_of = {0: 0.0, 1: 1.5707963, 2: 2 * 1.5707963, 3: -1.5707963}
    for i in rng.choice(n, 64, replace=False):
        e = oracle.XoEnv()
        lib.xo_env_init(C.byref(cfg), C.byref(e), 0)
        for c, v in enumerate(F["grid"][i]):
            e.grid[c] = int(v)
        e.agent_x, e.agent_y, e.agent_yaw = int(F["agent_x"][i]), int(F["agent_y"][i]), yaw_of[int(F["facing"][i])]
        for k in range(4):
            e.goal_icon[k] = int(F["goal_icon"][i, k])
            e.goal_yaw[k] = 0 + (1.5707963 * 4 - 0) * (int(F["goal_yaw"][i, k]) / 4096.0)
            e.goal_scale[k], e.goal_offset[k] = float(F["goal_scale"][i, k]), float(F["goal_offset"][i, k])
        want = np.zeros((3, 84, 84), np.uint8)
        lib.xo_render(C.byref(cfg), C.byref(synthetic_catalog.as_c()), C.byref(e), want.ctypes.data)
        assert (f1[i] == want).all(), i


def test_gray_frames(backend_cls, synthetic_catalog):
    """--color=false (the reference default): one plane, OpenCV 3.2.0's BGR2GRAY of the colour frame -- fully observed and
    first person, against the oracle."""
    for name in ("c3_nav2d_11x11_84", "fpv_nav3d_8x8_vr3_84", "c4_nav3d_15x15_128"):
        cfg = parity.make_cfg(name, gray=1)
        eng = backend_cls(cfg, synthetic_catalog, 256)
        orc = oracle.Oracle(cfg, synthetic_catalog, 256, threads=8)
        assert eng.sim.get_screen_out_dimensions()[2] == 1
        parity.run_parity(eng, orc, 24, render_every=6)


def test_invalid_action_status(backend_cls, synthetic_catalog):
    """an invalid action sets the env's error flag AND the call's status (the reference CHECK-aborts)."""
    import torch
    for name, n_act in (("c2_nav3d_7x7_84", 4), ("fpv_nav3d_8x8_vr3_84", 6)):
        cfg = parity.make_cfg(name)
        eng = backend_cls(cfg, synthetic_catalog, 64)
        eng.reset()
        sim = eng.sim
        assert sim.get_num_actions() == n_act
        a = torch.zeros(64, dtype=torch.int32, device="cuda")
        a[5] = n_act
        before = {k: eng.field(k) for k in ("grid", "agent_x", "agent_y", "num_steps")}
        sim._lib.xw_error_flags.argtypes = [C.c_void_p, C.c_void_p]
        rc = sim._lib.xw_step(sim._h, a.data_ptr(), 1, sim._d_reward.data_ptr(), sim._d_over.data_ptr(), None, sim._stream())
        torch.cuda.synchronize()
        assert sim._lib.xw_error_flags(sim._h, None) == 1
        flags = np.zeros(64, np.int32)
        assert sim._lib.xw_error_flags(sim._h, flags.ctypes.data) == 1 and flags[5] == _abi.XW_ERR_INVALID_ACTION and (np.delete(flags, 5) == 0).all()
        assert rc == 0   # the device-pointer call cannot know yet; the host-buffer calls return the status:
        for k in ("agent_x", "agent_y", "num_steps"):
            assert eng.field(k)[5] == before[k][5]   # the env was left untouched
        assert (eng.field("grid")[5] == before["grid"][5]).all()
        h_a = np.zeros(64, np.int32)
        h_a[7] = -3
        r, o = np.zeros(64, np.float32), np.zeros(64, np.int32)
        rc = sim._lib.xw_step_host(sim._h, h_a.ctypes.data, 1, r.ctypes.data, o.ctypes.data, None)
        assert rc == _abi.XW_ERR_INVALID_ACTION and b"invalid action" in sim._lib.xw_last_error()
        assert eng.field("num_steps")[5] == 1 and eng.field("num_steps")[7] == 1 and eng.field("num_steps")[0] == 2
        assert sim._lib.xw_error_flags(sim._h, None) == 2
        m = np.zeros(64, np.uint8)
        m[5] = 1
        eng.reset(m)   # a reset clears the env's flag
        assert sim._lib.xw_error_flags(sim._h, None) == 1


def test_debug_switches_are_compile_time_only(synthetic_catalog):
    """XW_RENDER_DEBUG used to switch parts of k_render_sp off at run time; the shipped build must ignore it (the switches
    exist only under -DXW_SP_DEBUG): frames stay oracle-exact with the variable set."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import oracle, parity\n"
            "from gpu_backend import EngineBackend\n"
            "from xworld_b200.catalog import Catalog\n"
            "cat = Catalog.synthetic(seed=0)\n"
            "cfg = parity.make_cfg('c3_nav2d_11x11_84')\n"
            "eng = EngineBackend(cfg, cat, 256); orc = oracle.Oracle(cfg, cat, 256, threads=4)\n"
            "assert eng.sim.render_kernel() == 3\n"
            "st = parity.run_parity(eng, orc, 12, render_every=4)\n"
            "print('frames', st['frames'])\n") % (parity.HERE + "/..", parity.HERE)
    env = dict(__import__("os").environ, XW_RENDER_DEBUG="29")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and "frames 1024" in out.stdout, out.stdout + out.stderr


def test_partial_batches_bookkeeping_and_context_history(synthetic_catalog):
    """game_over / lives of envs outside a reset mask stay what they were; an env that was reset and then sits a
    step out reads "alive"; the device-tensor path refreshes the host's view; with --context > 1 an env that sits a step out
    keeps its frame history and a masked / auto reset starts from a zero-filled context (simulator.cpp:110-113)."""
    import torch
    n = 64
    cfg = parity.make_cfg("c2_nav3d_7x7_84", max_steps=3, context=3)
    sim = Simulator("xworld", cfg, synthetic_catalog, n, -1)
    sim.reset_game()
    for s in range(3):
        sim.take_actions(np.full(n, s % 4, np.int32))
    assert all("max_step" in g for g in sim.game_over())
    mask = np.zeros(n, np.uint8)
    mask[::2] = 1
    before = sim.screen().cpu().numpy().reshape(n, 3, 3, 84, 84).copy()
    sim.reset_game(mask)
    go = sim.game_over()
    assert all(g == "alive" for g in go[::2]) and all("max_step" in g for g in go[1::2])
    assert (sim.get_lives()[::2] == 1).all() and (sim.get_lives()[1::2] == 0).all()
    scr = sim.screen().cpu().numpy().reshape(n, 3, 3, 84, 84)
    assert (scr[::2, :2] == 0).all() and (scr[::2, 2] != 0).any()        # reset envs: zero-filled history + the new frame
    assert (scr[1::2] == before[1::2]).all()                             # the others: history untouched by the re-render
    # the reset envs sit a step out, the others step: reset envs still read "alive" and keep their (zero) history
    a = np.where(mask == 1, _abi.XW_ACTION_NONE, 1).astype(np.int32)
    sim.take_actions(a)
    go = sim.game_over()
    assert all(g == "alive" for g in go[::2])
    scr2 = sim.screen().cpu().numpy().reshape(n, 3, 3, 84, 84)
    assert (scr2[::2] == scr[::2]).all()
    assert (scr2[1::2, 0] == scr[1::2, 1]).all() and (scr2[1::2, 1] == scr[1::2, 2]).all()   # stepped envs shifted by one
    # device-tensor path: game_over() follows the device-side step
    sim.reset_game()
    for s in range(3):
        sim.take_actions(torch.full((n,), s % 4, dtype=torch.int32, device="cuda"))
    assert all("max_step" in g for g in sim.game_over()) and (sim.get_lives() == 0).all()
    # get_state()'s extra keys (py_simulator.cpp:276-283) and the report
    one = Simulator("xworld", parity.make_cfg("ref_nav3d_8x8_96"), synthetic_catalog, 1, -1)
    one.reset_game()
    st = one.get_state()
    assert st["task"].startswith("XWorld3DNav") and st["event"] == "" and (st["height"], st["width"]) == ("8", "8")
    assert one.get_world_dimensions() == (8.0, 8.0, 0.0)
    lines, raw = one.teacher_report_task_performance()
    assert len(raw) == 5 and lines[0] == "=== XWorld3DNavTarget ==="
    w = Simulator("xworld", parity.make_cfg("ref_nav2d_8x8_96"), synthetic_catalog, 8, -1)
    w.reset_game()
    assert all(w.get_extra_info(e).split("task:")[1].startswith(("XWorldNav", "XWorldRec")) for e in range(8))
