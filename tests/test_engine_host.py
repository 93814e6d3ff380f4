"""CPU parity of the ENGINE's own per-env code: xworld_b200/csrc/*.cuh compiled for the host
(tests/hostsim) against the oracle -- reset, step/teacher, compose + straddle fix-up, race.
The same functions are what the CUDA kernels inline; the -m gpu tests repeat this through the C ABI."""
import ctypes as C

import numpy as np
import pytest

import oracle
import parity
from test_oracle_render import golden_cases
from xworld_b200 import _abi


@pytest.mark.parametrize("name", sorted(parity.CONFIGS))
def test_hostsim_matches_oracle(name, synthetic_catalog):
    cfg = parity.make_cfg(name)
    n = 48
    hs = parity.HostSim(cfg, synthetic_catalog, n)
    orc = oracle.Oracle(cfg, synthetic_catalog, n, threads=4)
    stats = parity.run_parity(hs, orc, 120, render_every=40)
    assert stats["frames"] >= 4 * n


@pytest.mark.parametrize("kw,n,steps", [
    (dict(curriculum=0.9, max_steps_factor=1), 8, 9000),                          # level 0 for ever: the 200-result windows fill and roll
    (dict(curriculum=0.02, start_level=2, curriculum_check_period=3), 16, 1500),  # offset-1 levels (displaced referent)
    (dict(curriculum=0.02, start_level=4, curriculum_check_period=2), 16, 1500),  # 7x7 world in the 8x8 map, then the last level
    (dict(curriculum=0.5, start_level=5), 16, 300),                               # last level = the curriculum-0 map
    (dict(curriculum=0.1, curriculum_check_period=4, auto_reset=1, max_steps=40), 48, 600),
])
def test_hostsim_curriculum(kw, n, steps, synthetic_catalog):
    """per-env level schedule, padded worlds, result windows -- engine code vs the oracle."""
    cfg = parity.make_cfg("curriculum_nav3d_8x8_96", **kw)
    hs = parity.HostSim(cfg, synthetic_catalog, n)
    orc = oracle.Oracle(cfg, synthetic_catalog, n, threads=2)
    parity.run_parity(hs, orc, steps, render_every=steps // 3, check_state_every=53, auto_reset=bool(kw.get("auto_reset")))
    lv = hs.field("level")
    assert lv.min() >= kw.get("start_level", 0)
    if kw.get("max_steps_factor") == 1:
        assert (hs.field("win_len") == 200).all()
    if kw.get("start_level") == 4:
        assert lv.max() == 5
    if kw.get("auto_reset"):
        assert lv.max() >= 1


def test_hostsim_action_none(synthetic_catalog):
    """XW_ACTION_NONE (include/xworld_b200.h): envs that sit a step out are untouched."""
    for name in ("c2_nav3d_7x7_84", "c3_nav2d_11x11_84"):
        cfg = parity.make_cfg(name)
        hs = parity.HostSim(cfg, synthetic_catalog, 24)
        orc = oracle.Oracle(cfg, synthetic_catalog, 24, threads=1)
        assert parity.run_partial_parity(hs, orc, 80, render_every=40) > 500


def test_hostsim_act_rep_and_global_ids(synthetic_catalog):
    cfg = parity.make_cfg("c2_nav3d_7x7_84", env_id_offset=1000, seed=42, simulator_seed=7)
    hs = parity.HostSim(cfg, synthetic_catalog, 32)
    orc = oracle.Oracle(cfg, synthetic_catalog, 32, threads=2)
    parity.run_parity(hs, orc, 60, act_rep=3)


def test_hostsim_auto_reset(synthetic_catalog):
    cfg = parity.make_cfg("c2_nav3d_7x7_84", auto_reset=1, max_steps=25)
    hs = parity.HostSim(cfg, synthetic_catalog, 40)
    orc = oracle.Oracle(cfg, synthetic_catalog, 40, threads=2)
    st = parity.run_parity(hs, orc, 80, render_every=20, auto_reset=True)
    assert st["events"].get(_abi.XW_MAX_STEP, 0) > 0


def test_hostsim_golden_frames_from_real_opencv():
    """Phase-atlas compositor vs frames the real OpenCV produced from the reference call sequence."""
    n = 0
    for tag, cfg, cat, grid, gi, want in golden_cases():
        hs = parity.HostSim(cfg, cat, 1)
        d_grid = np.zeros(cfg.height * cfg.width, np.uint8)
        # poke the state in: hostsim exposes the same SoA the kernels read
        L = hs.L
        L.hs_reset(hs.h, None)
        import ctypes
        # write grid + goal icons through the raw field pointers (test-only)
        out = hs.field("grid")
        assert out.shape[1] == len(grid)
        L.hs_set_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        g = np.ascontiguousarray(grid, np.uint8)
        icons = np.ascontiguousarray(gi, np.int32)
        L.hs_set_grid(hs.h, g.ctypes.data, icons.ctypes.data)
        got = hs.render()[0]
        assert (got == want).all(), (tag, int((got != want).sum()))
        n += 1
    assert n == 12


def test_hostsim_race_matches_oracle():
    lib = oracle.lib()
    for tt, full, hard, rep, ms, rnd in [(0, 0, 0, 1, 0, 0), (1, 1, 1, 1, 0, 0), (0, 1, 0, 3, 40, 0), (1, 1, 1, 4, 0, 0),
                                         (0, 1, 0, 1, 0, 1), (1, 1, 1, 2, 30, 1)]:   # (the last two: --random start states)
        cfg = _abi.default_config(game=_abi.XW_GAME_SIMPLE_RACE, track_type=tt, race_full_manouver=full, difficulty=hard,
                                  auto_reset=1, max_steps=ms, race_random=rnd, simulator_seed=5, env_id_offset=3)
        n = 64
        hs = parity.HostSim(cfg, None, n)
        hs.reset()
        orcs = [oracle.XoRace() for _ in range(n)]
        for i, o in enumerate(orcs):
            o.minstd = lib.xo_minstd_seed_for_thread(5, 3 + i + 1)
            lib.xo_race_reset(C.byref(cfg), C.byref(o))
        rng = np.random.RandomState(5)
        for s in range(200):
            a = rng.randint(0, 9 if full else 2, n).astype(np.int32)
            r, ov, _ = hs.step(a, act_rep=rep)
            st = hs.field("state")
            for i, o in enumerate(orcs):
                st2, ov2 = (C.c_float * 4)(), C.c_int32()
                r2 = lib.xo_race_take_actions(C.byref(cfg), C.byref(o), int(a[i]), rep, st2, C.byref(ov2))
                assert np.float32(r2).view(np.uint32) == r[i].view(np.uint32) and ov2.value == ov[i], (tt, rep, s, i)
                assert (np.array(list(st2), np.float32).view(np.uint32) == st[i].view(np.uint32)).all()
                if ov2.value:
                    lib.xo_race_reset(C.byref(cfg), C.byref(o))


@pytest.mark.parametrize("geom", [(11, 11, 84, 84), (7, 7, 84, 84), (15, 15, 128, 128), (8, 8, 96, 96), (16, 16, 128, 128),
                                  (10, 10, 84, 84), (12, 12, 100, 100), (16, 16, 84, 84), (6, 6, 64, 64), (13, 13, 128, 128),
                                  (5, 9, 60, 100), (16, 16, 252, 252), (3, 3, 20, 20), (9, 9, 84, 84), (14, 14, 112, 112)])
def test_sparse_painter_and_plan_compositor_vs_per_pixel_rule(geom, synthetic_catalog):
    """Both shared-memory renderers against the per-pixel rule on crowded random grids: goals and the agent
    packed next to each other and next to bricks, so that every pair-table / exact fallback branch runs
    (special-special borders, corners of four different cells)."""
    H, W, OH, OW = geom
    G = 8
    cfg = _abi.default_config(height=H, width=W, n_goals=G, n_blocks=1, rules=_abi.XW_RULES_NAV3D, out_h=OH, out_w=OW)
    n = 24
    hs = parity.HostSim(cfg, synthetic_catalog, n)   # no reset: the grids are written below (maps are square
    L = hs.L                                          # in the engine; the renderers take any H x W)
    L.hs_set_grid_env.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.RandomState(H * 1000 + OW)
    pool = rng.permutation(synthetic_catalog.as_c().n_icons)[:10]
    for e in range(n):
        grid = np.zeros(H * W, np.uint8)
        density = [0.0, 0.1, 0.3, 0.6, 1.0][e % 5]
        grid[rng.rand(H * W) < density] = _abi.XW_CELL_BLOCK
        # the agent and the goals in one clump (e even) or anywhere (e odd)
        k = min(G + 1, H * W)
        if e % 2 == 0:
            y0, x0 = rng.randint(0, max(1, H - 2)), rng.randint(0, max(1, W - 2))
            cand = [(y0 + dy) * W + (x0 + dx) for dy in range(3) for dx in range(3) if y0 + dy < H and x0 + dx < W]
            cells = rng.permutation(cand)[:k]
        else:
            cells = rng.permutation(H * W)[:k]
        for i, c in enumerate(cells):
            grid[c] = _abi.XW_CELL_AGENT + i
        icons = pool[rng.randint(0, len(pool), G)].astype(np.int32)
        L.hs_set_grid_env(hs.h, e, np.ascontiguousarray(grid).ctypes.data, icons.ctypes.data)
    want = hs.render(mode=2)
    if L.hs_fast_ok(hs.h):
        got0 = hs.render(mode=0)
        assert (got0 == want).all(), ("plan", geom, int((got0 != want).sum()))
    if L.hs_sp_ok(hs.h):
        got1 = hs.render(mode=1)
        bad = np.argwhere(got1 != want)
        assert bad.size == 0, ("sparse", geom, len(bad), bad[:8].tolist())
    assert L.hs_sp_ok(hs.h) == L.hs_fast_ok(hs.h)


@pytest.mark.parametrize("geom", [(11, 11, 84, 84), (15, 15, 128, 128), (16, 16, 128, 128), (10, 10, 84, 84), (5, 9, 60, 100), (13, 13, 128, 128)])
def test_sparse_painter_merge_mode_forced(geom, synthetic_catalog, monkeypatch):
    """The class tables' merge mode (XwRender::ctab_merge: brick|white words of shared word columns without a straddling pixel are
    the brick|brick word with the white cell's bytes set) is chosen only where it buys a third frame buffer (15x15 -> 128x128);
    forced on (XW_RENDER_CTAB_MERGE=1) it must stay pixel-exact on every geometry, maze-like grids included."""
    monkeypatch.setenv("XW_RENDER_CTAB_MERGE", "1")
    H, W, OH, OW = geom
    cfg = _abi.default_config(height=H, width=W, n_goals=4, n_blocks=1, rules=_abi.XW_RULES_NAV3D, out_h=OH, out_w=OW)
    n = 20
    hs = parity.HostSim(cfg, synthetic_catalog, n)
    L = hs.L
    L.hs_set_grid_env.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    if not L.hs_sp_ok(hs.h):
        pytest.skip("the painter does not take this geometry")
    rng = np.random.RandomState(H * 77 + OW)
    pool = rng.permutation(synthetic_catalog.as_c().n_icons)[:8]
    for e in range(n):
        grid = np.zeros(H * W, np.uint8)
        grid[rng.rand(H * W) < [0.15, 0.3, 0.5, 0.8][e % 4]] = _abi.XW_CELL_BLOCK   # brick | white and white | brick borders of every kind
        for i, c in enumerate(rng.permutation(H * W)[:5]):
            grid[c] = _abi.XW_CELL_AGENT + i
        icons = pool[rng.randint(0, len(pool), 4)].astype(np.int32)
        L.hs_set_grid_env(hs.h, e, np.ascontiguousarray(grid).ctypes.data, icons.ctypes.data)
    want = hs.render(mode=2)
    got = hs.render(mode=1)
    bad = np.argwhere(got != want)
    assert bad.size == 0, ("merge mode", geom, len(bad), bad[:8].tolist())


def test_hostsim_fpv_golden_frames_from_real_opencv():
    """The engine's first-person code (goal icon warp, crop-cell classes, table pass + exact pass) vs frames the real OpenCV
    produced from the reference call sequence (tests/golden/fpv_golden.npz)."""
    from test_oracle_fpv import fpv_golden_cases
    n = 0
    for tag, cfg, cat, grid, gi, pose, agent_yaw, want in fpv_golden_cases():
        hs = parity.HostSim(cfg, cat, 1)
        yaw_idx = np.array([int(round(p[0] / (1.5707963 * 4) * 4096)) for p in pose], np.uint16)
        scale = np.ascontiguousarray([p[1] for p in pose], np.float64)
        offset = np.ascontiguousarray([p[2] for p in pose], np.float64)
        facing = oracle.lib().xo_facing_dir(agent_yaw)
        hs.L.hs_set_fpv_env(hs.h, 0, np.ascontiguousarray(grid, np.uint8).ctypes.data, np.ascontiguousarray(gi, np.int32).ctypes.data,
                            yaw_idx.ctypes.data, scale.ctypes.data, offset.ctypes.data, facing)
        for mode in (1, 2):
            got = hs.render(mode=mode)[0]
            assert (got == want).all(), (tag, n, mode, int((got != want).sum()))
        n += 1
    assert n == 42
