"""The C-ABI shared library: loads, exports every symbol include/xworld_b200.h declares, refuses
to run the CUDA games without a device (no CPU fallback), and serves BASELINE config 1
(SimpleGame, batch 1, CPU plumbing through the SimulatorInterface-shaped API)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from xworld_b200 import Simulator, _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    return _abi.load()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "xworld_b200.h")).read()
    declared = set(re.findall(r"\b(xw_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"xw_sim"}
    assert declared == set(_abi.SYMBOLS), declared ^ set(_abi.SYMBOLS)
    for s in declared:
        assert getattr(lib, s) is not None


def test_config_struct_matches_header(lib):
    cfg = _abi.XwConfig()
    lib.xw_config_init(C.byref(cfg))
    ref = _abi.default_config()
    for f, _t in _abi.XwConfig._fields_:
        if f == "reserved":
            continue
        assert getattr(cfg, f) == getattr(ref, f), f


def test_no_cpu_fallback_for_cuda_games(lib, synthetic_catalog):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    cfg = _abi.default_config(height=7, width=7, n_goals=4, n_blocks=12)
    rc = lib.xw_create(C.byref(cfg), C.byref(synthetic_catalog.as_c()), 4, -1, C.byref(h))
    assert rc == -3 and b"no CPU fallback" in lib.xw_last_error()
    with pytest.raises(RuntimeError):
        Simulator.create("simple_race", {"track_width": 20.0, "track_length": 100.0, "track_radius": 30.0})


def test_invalid_arguments_return_errors_not_aborts(lib):
    h = C.c_void_p()
    cfg = _abi.default_config()
    cfg.abi_version = 99
    assert lib.xw_create(C.byref(cfg), None, 1, -1, C.byref(h)) == -1
    cfg = _abi.default_config(game=_abi.XW_GAME_SIMPLE_GAME, array_size=1)
    assert lib.xw_create(C.byref(cfg), None, 1, -1, C.byref(h)) == -1
    with pytest.raises(RuntimeError):
        Simulator.create("tetris", {})
    with pytest.raises(RuntimeError):
        Simulator.create("simple_game", {})  # array_size is a required key (py_simulator.cpp:99)


def test_config1_simple_game_known_answer(lib):
    """tests/test_simple_game_simulator.cpp:21-47 replayed through the reference's Python API names."""
    sim = Simulator.create("simple_game", {"array_size": 8})
    assert sim.get_num_actions() == 2 and sim.get_screen_out_dimensions() == [1, 8, 1, 1]
    sim.reset_game()
    pos = 4
    for i in range(3):
        scr = sim.get_state()["screen"]
        assert [round(v * 255) for v in scr] == [1 if j == pos else 0 for j in range(8)]
        r = sim.take_action({"action": 1})
        pos += 1
        assert abs(r - (2.0 if pos == 7 else -0.1)) < 1e-6
    assert sim.game_over() == "success" and sim.get_num_steps() == 3 and sim.get_lives() == 0
    sim.reset_game()
    assert sim.game_over() == "alive"
    assert abs(sim.take_actions({"action": 0}, 4, False) - (-0.1 * 3 + 4.0)) < 1e-6  # act_rep sums rewards


def test_simple_game_batch_vs_oracle(lib, oracle_lib):
    import oracle
    n = 16
    sim = Simulator.create("simple_game", {"array_size": 9, "n_envs": n})
    games = [oracle.XoSimpleGame() for _ in range(n)]
    for g in games:
        oracle_lib.xo_sg_reset(C.byref(g), 9)
    rng = np.random.RandomState(0)
    for s in range(12):
        a = rng.randint(0, 2, n)
        r = sim.take_actions(a)
        for i, g in enumerate(games):
            assert r[i] == oracle_lib.xo_sg_act(C.byref(g), int(a[i]))
            assert (sim.screen()[i] == np.array(list(g.state)[:9], np.uint8)).all()
