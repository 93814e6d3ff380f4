"""Pins the oracle's renderer: cv::resize arithmetic against the live cv2 (when importable) and
against frames the real OpenCV produced from the reference's call sequence (render_golden.npz)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from xworld_b200 import _abi
from xworld_b200.catalog import Catalog

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_catalog():
    z = np.load(os.path.join(HERE, "golden", "render_golden.npz"))
    metas = []
    for p in z["paths"]:
        p = str(p)
        parts = p.split("/")
        metas.append({"path": p, "type": parts[0], "name": "_".join(os.path.basename(p).split("_")[:-1]),
                      "subtree": parts[1] if parts[0] == "goal" else "", "color": "na"})
    order = sorted(range(len(metas)), key=lambda i: metas[i]["path"])
    remap = {old: new for new, old in enumerate(order)}
    cat = Catalog([metas[i] for i in order], z["atlas"][order])
    return z, cat, remap


GOLDEN_CFG = {"c2": (7, 0), "c3": (11, 84), "c4": (15, 128), "ref8": (8, 0)}


def golden_cases():
    z, cat, remap = golden_catalog()
    for tag, (H, out) in GOLDEN_CFG.items():
        cfg = _abi.default_config(height=H, width=H, n_goals=4, n_blocks=1, out_h=out, out_w=out)
        for i in range(len(z[tag + "_grid"])):
            gi = [remap[int(v)] for v in z[tag + "_goal_icon"][i]]
            yield tag, cfg, cat, z[tag + "_grid"][i], gi, z[tag + "_frames"][i]


def test_golden_frames_from_real_opencv(oracle_lib):
    n = 0
    for tag, cfg, cat, grid, gi, want in golden_cases():
        e = oracle.XoEnv()
        oracle_lib.xo_env_init(C.byref(cfg), C.byref(e), 0)
        for c, v in enumerate(grid):
            e.grid[c] = int(v)
        for k in range(4):
            e.goal_icon[k] = gi[k]
        out = np.zeros_like(want)
        oracle_lib.xo_render(C.byref(cfg), C.byref(cat.as_c()), C.byref(e), out.ctypes.data)
        assert (out == want).all(), (tag, int((out != want).sum()))
        n += 1
    assert n == 12


def test_resize_tables_and_pixels_vs_live_cv2(oracle_lib):
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    rng = np.random.RandomState(0)
    for s, d in [(704, 84), (960, 128), (448, 84), (512, 96), (192, 84), (320, 80), (64, 12), (128, 100)]:
        img = rng.randint(0, 256, (s, s, 3)).astype(np.uint8)
        ref = cv2.resize(img, (d, d), interpolation=cv2.INTER_LINEAR)
        out = np.zeros((d, d, 3), np.uint8)
        oracle_lib.xo_resize_linear_8uc3(img.ctypes.data, s, s, out.ctypes.data, d, d)
        assert (out == ref).all(), (s, d, int(np.abs(out.astype(int) - ref).max()))
    # non-square
    img = rng.randint(0, 256, (448, 704, 3)).astype(np.uint8)
    ref = cv2.resize(img, (90, 70), interpolation=cv2.INTER_LINEAR)
    out = np.zeros((70, 90, 3), np.uint8)
    oracle_lib.xo_resize_linear_8uc3(img.ctypes.data, 448, 704, out.ctypes.data, 70, 90)
    assert (out == ref).all()


def test_tile_periodicity_claim(oracle_lib, synthetic_catalog):
    """at 7x7 -> 84x84 resizing the canvas == resizing each 64-px tile to 12 px."""
    cfg = _abi.default_config(height=7, width=7, n_goals=4, n_blocks=12, seed=3, simulator_seed=1)
    o = oracle.Oracle(cfg, synthetic_catalog, 4)
    o.reset()
    frames = o.render()
    for i, e in enumerate(o.envs):
        for cy in range(7):
            for cx in range(7):
                code = e.grid[cy * 7 + cx]
                tile = frames[i, :, cy * 12:(cy + 1) * 12, cx * 12:(cx + 1) * 12]
                if code == 0:
                    assert (tile == 255).all()
                    continue
                icon = synthetic_catalog.brick_icon if code == 1 else synthetic_catalog.agent_icon if code == 2 \
                    else e.goal_icon[code - 3]
                small = np.zeros((12, 12, 3), np.uint8)
                src = np.ascontiguousarray(synthetic_catalog.atlas64[icon])
                oracle_lib.xo_resize_linear_8uc3(src.ctypes.data, 64, 64, small.ctypes.data, 12, 12)
                assert (tile == small.transpose(2, 0, 1)).all()
