"""Pins the oracle's first-person view (--visible_radius > 0, oracle/xw_oracle_fpv.c):
  * cv::getRotationMatrix2D / cv::warpAffine against the live cv2 (when importable),
  * whole frames against what the real OpenCV produced from the reference's call sequence, with the ROI / shadow flags of
    the reference's own compiled XMap::image_masking (tests/golden/gen_fpv_golden.py -> fpv_golden.npz),
  * image_masking and the six-action XAgent::act against the reference's compiled xmap.cpp / xitem.cpp (oracle/_ref),
  * the reset-time pose draws and the step / teacher rules against the reference's own Python run with
    --visible_radius > 0 (tests/golden/gen_reference_python.py fpv -> refpy_fpv.json.gz)."""
import ctypes as C
import gzip
import json
import math
import os

import numpy as np
import pytest

import oracle
from xworld_b200 import _abi
from xworld_b200.catalog import Catalog

HERE = os.path.dirname(os.path.abspath(__file__))
PI_2 = 1.5707963

FPV_CASES = {"vr3_7x7": (7, 3), "vr7_11x11": (11, 7), "vr5_8x8": (8, 5), "vr7_7x7": (7, 7), "vr1_7x7": (7, 1),
             "vr9_15x15": (15, 9), "vr11_11x11": (11, 11)}


def fpv_golden_catalog():
    z = np.load(os.path.join(HERE, "golden", "fpv_golden.npz"))
    metas = []
    for p in z["paths"]:
        p = str(p)
        parts = p.split("/")
        metas.append({"path": p, "type": parts[0], "name": "_".join(os.path.basename(p).split("_")[:-1]),
                      "subtree": parts[1] if parts[0] == "goal" else "", "color": "na"})
    order = sorted(range(len(metas)), key=lambda i: metas[i]["path"])
    remap = {old: new for new, old in enumerate(order)}
    return z, Catalog([metas[i] for i in order], z["atlas"][order]), remap


def fpv_golden_cases():
    """(tag, cfg, catalog, grid, goal icons, goal poses [4][yaw, scale, offset], agent yaw, frame)"""
    z, cat, remap = fpv_golden_catalog()
    for tag, (H, vr) in FPV_CASES.items():
        cfg = _abi.default_config(height=H, width=H, n_goals=4, n_blocks=1, visible_radius=vr)
        for i in range(len(z[tag + "_grid"])):
            gi = [remap[int(v)] for v in z[tag + "_goal_icon"][i]]
            yield tag, cfg, cat, z[tag + "_grid"][i], gi, z[tag + "_goal_pose"][i], float(z[tag + "_agent_yaw"][i]), z[tag + "_frames"][i]


def load_env(oracle_lib, cfg, grid, gi, pose, agent_yaw):
    e = oracle.XoEnv()
    oracle_lib.xo_env_init(C.byref(cfg), C.byref(e), 0)
    for c, v in enumerate(grid):
        e.grid[c] = int(v)
        if v == _abi.XW_CELL_AGENT:
            e.agent_x, e.agent_y = c % cfg.width, c // cfg.width
    for k in range(4):
        e.goal_icon[k] = gi[k]
        e.goal_yaw[k], e.goal_scale[k], e.goal_offset[k] = [float(v) for v in pose[k]]
    e.agent_yaw = agent_yaw
    return e


def test_fpv_golden_frames_from_real_opencv(oracle_lib):
    n = 0
    for tag, cfg, cat, grid, gi, pose, agent_yaw, want in fpv_golden_cases():
        e = load_env(oracle_lib, cfg, grid, gi, pose, agent_yaw)
        out = np.zeros_like(want)
        oracle_lib.xo_render(C.byref(cfg), C.byref(cat.as_c()), C.byref(e), out.ctypes.data)
        assert (out == want).all(), (tag, n, int((out != want).sum()))
        n += 1
    assert n == 6 * len(FPV_CASES)


def test_warp_affine_vs_live_cv2(oracle_lib):
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    cv2.ipp.setUseIPP(False)
    rng = np.random.RandomState(5)
    white = np.array([255, 255, 255], np.uint8)
    for t in range(200):
        icon = rng.randint(0, 256, (64, 64, 3)).astype(np.uint8)
        yaw, scale = rng.uniform(0, PI_2 * 4), rng.uniform(0.5, 1)
        if t < 8:  # the agent's four headings, exact and drifted by turns
            yaw, scale = [-PI_2, 0.0, PI_2, 2 * PI_2, math.pi / 2, math.pi, -math.pi / 2, 3.1415926 - 2 * math.pi + 2 * math.pi][t], 1.0
        offset = rng.uniform(0, 1 - scale) if scale < 1 else 0.0
        M = cv2.getRotationMatrix2D((32.0, 32.0), 90 - yaw * 180 / math.pi, scale)
        Mo = (C.c_double * 6)()
        oracle_lib.xo_rotation_matrix(32.0, 32.0, 90 - yaw * 180 / math.pi, scale, Mo)
        assert list(Mo) == list(M.ravel()), t
        M[0, 2] += (offset + scale / 2 - 0.5) * 64
        M[1, 2] += (offset + scale / 2 - 0.5) * 64
        want = cv2.warpAffine(icon, M, (64, 64), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=(255, 255, 255))
        got = np.zeros_like(icon)
        Mc = (C.c_double * 6)(*M.ravel())
        oracle_lib.xo_warp_affine_8uc3(icon.ctypes.data, 64, 64, got.ctypes.data, 64, 64, Mc, white.ctypes.data)
        assert (got == want).all(), (t, yaw, scale, offset, int((got != want).sum()))
        got2 = np.zeros_like(icon)
        oracle_lib.xo_item_image(icon.ctypes.data, yaw, scale, offset, got2.ctypes.data)
        assert (got2 == want).all(), t
    # the view rotation (xmap.cpp:196-200): 90 + yaw about the view centre, black border, all four headings and sizes
    black = np.zeros(3, np.uint8)
    for vs in (64, 192, 448):
        view = rng.randint(0, 256, (vs, vs, 3)).astype(np.uint8)
        for yaw in (-PI_2, 0.0, PI_2, 2 * PI_2, math.pi, -math.pi / 2 - 1e-9):
            M = cv2.getRotationMatrix2D((vs / 2.0, vs / 2.0), 90 + yaw * 180 / math.pi, 1.0)
            want = cv2.warpAffine(view, M, (vs, vs))
            got = np.zeros_like(view)
            oracle_lib.xo_warp_affine_8uc3(view.ctypes.data, vs, vs, got.ctypes.data, vs, vs, (C.c_double * 6)(*M.ravel()), black.ctypes.data)
            assert (got == want).all(), (vs, yaw)


def test_view_rotation_is_a_pixel_permutation(oracle_lib):
    """What the engine relies on: for the agent's four headings warpAffine(90 + yaw) about (N/2, N/2) is an
    exact pixel permutation -- a quarter-turn rotation about the pixel CORNER (N/2, N/2), which maps index i to N - i and so
    leaves one black row / column: dst(y, x) = src(sy, sx) with
        up: (y, x)   right: (x, N - y)   down: (N - y, N - x)   left: (N - x, y),     black where an index equals N.
    Also with the drifted doubles turning produces (xitem.cpp:140-151)."""
    rng = np.random.RandomState(2)
    black = np.zeros(3, np.uint8)
    for N in (64, 192, 448):
        view = rng.randint(1, 256, (N, N, 3)).astype(np.uint8)
        padded = np.zeros((N + 1, N + 1, 3), np.uint8)
        padded[:N, :N] = view
        yy, xx = np.mgrid[0:N, 0:N]
        maps = {3: (yy, xx), 0: (xx, N - yy), 1: (N - yy, N - xx), 2: (N - xx, yy)}
        for yaw in (0.0, PI_2, 2 * PI_2, -PI_2, math.pi / 2, math.pi, -math.pi / 2, 3.1415926 + math.pi / 2 - 2 * math.pi):
            M = (C.c_double * 6)()
            oracle_lib.xo_rotation_matrix(N / 2.0, N / 2.0, 90 + yaw * 180 / math.pi, 1.0, M)
            got = np.zeros_like(view)
            oracle_lib.xo_warp_affine_8uc3(view.ctypes.data, N, N, got.ctypes.data, N, N, M, black.ctypes.data)
            sy, sx = maps[oracle_lib.xo_facing_dir(yaw)]
            assert (got == padded[sy, sx]).all(), (N, yaw)


def test_agent_icon_rotation_is_a_pixel_permutation(oracle_lib):
    """The agent's icon (xitem.cpp:47-60 with the agent's yaw, scale 1, offset 0): rotation by 90 - yaw about (32, 32), white
    border -- quarter turns again: dst(y, x) = src(sy, sx) with down: (y, x), right: (x, 64 - y), left: (64 - x, y),
    up: (64 - y, 64 - x); white where an index equals 64."""
    rng = np.random.RandomState(3)
    icon = rng.randint(0, 255, (64, 64, 3)).astype(np.uint8)
    padded = np.full((65, 65, 3), 255, np.uint8)
    padded[:64, :64] = icon
    yy, xx = np.mgrid[0:64, 0:64]
    maps = {1: (yy, xx), 0: (xx, 64 - yy), 2: (64 - xx, yy), 3: (64 - yy, 64 - xx)}
    for yaw in (0.0, PI_2, 2 * PI_2, -PI_2, math.pi / 2, math.pi, -math.pi / 2):
        got = np.zeros_like(icon)
        oracle_lib.xo_item_image(icon.ctypes.data, yaw, 1.0, 0.0, got.ctypes.data)
        sy, sx = maps[oracle_lib.xo_facing_dir(yaw)]
        assert (got == padded[sy, sx]).all(), yaw


def test_gray_formula(oracle_lib, synthetic_catalog):
    """--color=false: the frame is cvtColor(BGR2GRAY) of the colour frame with OpenCV 3.2.0's 14-bit coefficients; the same
    structure with OpenCV 4's 15-bit coefficients equals the live cv2 (so only the constants are version-bound)."""
    cfg = _abi.default_config(height=7, width=7, n_goals=4, n_blocks=12, seed=3, simulator_seed=1)
    o = oracle.Oracle(cfg, synthetic_catalog, 3)
    o.reset()
    col = o.render()
    cfg_g = _abi.default_config(height=7, width=7, n_goals=4, n_blocks=12, seed=3, simulator_seed=1, gray=1)
    og = oracle.Oracle(cfg_g, synthetic_catalog, 3)
    og.reset()
    g = og.render()
    assert g.shape == (3, 1, 84, 84)
    b, gr, r = [col[:, c].astype(np.int64) for c in range(3)]
    assert (g[:, 0] == ((b * 1868 + gr * 9617 + r * 4899 + 8192) >> 14)).all()
    try:
        import cv2
    except ImportError:
        return
    hwc = np.ascontiguousarray(col[0].transpose(1, 2, 0))
    want = cv2.cvtColor(hwc, cv2.COLOR_BGR2GRAY)
    assert (want == ((b[0] * 3735 + gr[0] * 19235 + r[0] * 9798 + 16384) >> 15)).all()


def _ref():
    if not os.path.exists(oracle.REF_LIB):
        pytest.skip("oracle/_ref/libxw_ref.so not built (needs /root/reference)")
    R = C.CDLL(oracle.REF_LIB)
    R.ref_map_create.restype = C.c_void_p
    R.ref_map_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int]
    R.ref_map_act.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
    R.ref_map_destroy.argtypes = [C.c_void_p]
    R.ref_map_masking.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    R.ref_map_num_actions.argtypes = [C.c_void_p]
    return R


def _ref_map(R, e, D, vr):
    grid = np.array(list(e.grid)[:D * D])
    cells = np.nonzero(grid)[0]
    types = np.array([0 if grid[c] == 1 else 2 if grid[c] == 2 else 1 for c in cells], np.int32)
    xs, ys = (cells % D).astype(np.int32), (cells // D).astype(np.int32)
    return R.ref_map_create(D, D, len(cells), types.ctypes.data, xs.ctypes.data, ys.ctypes.data, e.agent_yaw, vr), grid, cells


@pytest.mark.parametrize("D,vr", [(7, 3), (11, 7), (8, 5), (7, 7), (11, 1), (15, 9)])
def test_six_actions_and_masking_vs_compiled_reference(oracle_lib, synthetic_catalog, D, vr):
    """XAgent::act with the FPV action set (xitem.cpp:89-155, incl. turns: yaw arithmetic, move_item returning false) and
    XMap::image_masking (xmap.cpp:273-362), compiled from the reference, against the oracle, step by step."""
    R = _ref()
    blocks = {7: 12, 8: 16, 11: 30, 15: 56}[D]
    cfg = _abi.default_config(height=D, width=D, n_goals=4, n_blocks=blocks, rules=0, seed=91, simulator_seed=2, visible_radius=vr)
    o = oracle.Oracle(cfg, synthetic_catalog, 12)
    o.reset()
    rng = np.random.RandomState(D * 31 + vr)
    headings = set()
    for e in o.envs:
        m, grid, cells = _ref_map(R, e, D, vr)
        assert R.ref_map_num_actions(m) == 6
        for s in range(80):
            rect = np.zeros(4, np.int32)
            sh_ref = np.zeros(vr * vr, np.uint8)
            R.ref_map_masking(m, vr, rect.ctypes.data, sh_ref.ctypes.data)
            xs, ys = C.c_int(), C.c_int()
            sh = np.zeros(vr * vr, np.uint8)
            oracle_lib.xo_image_masking(C.byref(e), vr, C.byref(xs), C.byref(ys), sh.ctypes.data)
            assert (xs.value, ys.value, vr, vr) == tuple(rect) and (sh == sh_ref).all(), (s, e.agent_yaw)
            headings.add(oracle_lib.xo_facing_dir(e.agent_yaw))
            a = int(rng.randint(0, 6))
            ax, ay, nc, ct = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            yaw = C.c_double()
            ok = R.ref_map_act(m, a, C.byref(ax), C.byref(ay), C.byref(yaw), C.byref(ct), C.byref(nc))
            r, ov = C.c_float(), C.c_int32()
            before = np.array(list(e.grid)[:D * D])
            rc = oracle_lib.xo_step(C.byref(cfg), C.byref(o.cat_c), C.byref(e), a, 1, C.byref(r), C.byref(ov))
            assert rc == 0
            assert (e.agent_x, e.agent_y, e.action_success) == (ax.value, ay.value, ok), (s, a)
            assert e.agent_yaw == yaw.value, (s, a)                 # bit-equal doubles, turns drift identically
            if a >= 4:
                assert ok == 0 and nc.value == 0                   # the "turn returns false" quirk
            if nc.value:
                assert before[cells[ct.value]] != 0
            if e.stage == _abi.XW_STAGE_TERMINAL:
                break
        R.ref_map_destroy(m)
    assert headings == {0, 1, 2, 3}


def test_goal_pose_draws(oracle_lib):
    """uniform(a, b) = a + (b - a) * random(): the ranges of xworld_env.py:211-223 and the yaw grid."""
    for k in range(200):
        y, s, o = C.c_double(), C.c_double(), C.c_double()
        oracle_lib.xo_goal_pose(1234, k, 1, 0, k % 4, C.byref(y), C.byref(s), C.byref(o))
        assert 0 <= y.value < PI_2 * 4 and 0.5 <= s.value < 1 and 0 <= o.value < 1 - s.value
        idx = y.value / (PI_2 * 4) * oracle.YAW_STEPS
        assert abs(idx - round(idx)) < 1e-9


EVMAP = {"": 0, "correct_goal": 1, "wrong_goal": 2, "time_up": 3}
# XWorldSimulator::game_over (xworld_simulator.cpp:165-198): lang_acquisition maps the teacher's event; one_channel never ends
OVER_OF_EVENT = {"": 0, "correct_goal": _abi.XW_SUCCESS, "wrong_goal": _abi.XW_DEAD, "time_up": _abi.XW_MAX_STEP}


def test_fpv_and_one_channel_against_reference_python(oracle_lib, synthetic_catalog):
    """The reference's own Python (map generator with set_property's yaw / scale / offset draws, the task classes) run
    with --visible_radius > 0 and / or --task_mode=one_channel by tests/golden/gen_reference_python.py fpv: every reset
    (map, poses as bit-equal doubles, task bookkeeping) and every step (reward bits, position, yaw, success flag, event,
    game_over) of the oracle."""
    with gzip.open(os.path.join(HERE, "golden", "refpy_fpv.json.gz")) as f:
        tr = json.loads(f.read().decode())
    n_steps = n_resets = n_turns = n_events = 0
    for case in tr["cases"]:
        D, G = case["dim"], case["n_goals"]
        one = case["task_mode"] == "one_channel"
        cfg = _abi.default_config(height=D, width=D, n_goals=G, n_blocks=case["n_blocks"], rules=case["rules"], seed=case["seed"],
                                  simulator_seed=case["simulator_seed"], visible_radius=case["visible_radius"],
                                  task_mode=_abi.XW_TASK_ONE_CHANNEL if one else _abi.XW_TASK_LANG_ACQUISITION)
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
                assert list(e.goal_x)[:G] == rs["goal_x"] and list(e.goal_y)[:G] == rs["goal_y"], where
                assert list(e.goal_name)[:G] == rs["goal_name"] and list(e.goal_icon)[:G] == rs["goal_icon"], where
                assert e.agent_yaw == rs["agent_yaw"], where
                assert [[e.goal_yaw[k], e.goal_scale[k], e.goal_offset[k]] for k in range(G)] == rs["goal_pose"], where
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
                    n_turns += s["a"] >= 4
                    n_events += s["ev"] != ""
                    assert np.float32(r.value).view(np.uint32) == np.float32(s["r"]).view(np.uint32), (where, i)
                    assert [e.agent_x, e.agent_y] == s["agent"] and e.action_success == s["ok"], (where, i)
                    assert e.agent_yaw == s["yaw"], (where, i)
                    assert e.event == EVMAP[s["ev"]], (where, i)
                    assert ov.value == (0 if one else OVER_OF_EVENT[s["ev"]]), (where, i)
                    if case["rules"] == 1:
                        assert (e.stage != 0) == (s["stage"] != "idle") and e.task == s["task"], (where, i)
                assert e.minstd == ep["minstd"], where
    assert n_steps > 9000 and n_resets > 100 and n_turns > 1500 and n_events > 50
