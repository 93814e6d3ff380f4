#!/usr/bin/env python
"""Tuning sweep of the render kernel's warp-group shape (XW_RENDER_GROUPS / XW_RENDER_GROUP_THREADS /
XW_RENDER_SPLIT_M3) on one GPU: CUDA-event time of xw_render at the bench workload, plus a pixel check
against a reference frame set.  Usage: python tools/sweep_render.py [c3|c4|c2] [n_envs]"""
import ctypes as C
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from xworld_b200 import _abi
from xworld_b200.catalog import Catalog
from xworld_b200.simulator import Simulator

WL = {"c3": dict(height=11, width=11, n_goals=4, n_blocks=30, rules=1, out_h=84, out_w=84, max_steps=242, auto_reset=1),
      "c4": dict(height=15, width=15, n_goals=4, n_blocks=56, rules=0, out_h=128, out_w=128, auto_reset=1),
      "c2": dict(height=7, width=7, n_goals=4, n_blocks=12, rules=0, auto_reset=1)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    combos = json.loads(sys.argv[3]) if len(sys.argv) > 3 else None
    cat = Catalog.synthetic(seed=0)
    cfg = _abi.default_config(seed=1234, simulator_seed=1, **WL[name])
    fb = 3 * (cfg.out_h or cfg.height * 12) * (cfg.out_w or cfg.width * 12)
    ref = None
    if combos is None:
        combos = [(g, t, s) for g in (8, 7, 6, 5, 4, 3, 2) for t in (64, 96, 128, 160, 192, 256, 320) for s in (0, 1) if g * t <= 1024]
    for combo in combos:
        g, t, s = combo[:3]
        os.environ["XW_RENDER_CONFLICT_FREE"] = str(combo[3] if len(combo) > 3 else 0)
        os.environ["XW_RENDER_MODE"] = str(combo[4]) if len(combo) > 4 else "sb"
        os.environ["XW_RENDER_TWO_PHASE"] = str(combo[5]) if len(combo) > 5 else "0"
        os.environ["XW_RENDER_SP_FILL"] = str(combo[6]) if len(combo) > 6 else "0"
        os.environ["XW_RENDER_GROUPS"] = str(g)
        os.environ["XW_RENDER_GROUP_THREADS"] = str(t)
        os.environ["XW_RENDER_SPLIT_M3"] = str(s)
        try:
            sim = Simulator("xworld", cfg, cat, n, 0)
        except RuntimeError as e:
            print(json.dumps({"G": g, "GT": t, "split": s, "error": str(e)[:80]}))
            continue
        sim.reset_game()
        lib, h = sim._lib, sim._h
        scr = sim.screen()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            lib.xw_render(h, scr.data_ptr(), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 20
        for _ in range(K):
            lib.xw_render(h, scr.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        chk = scr[:2048].cpu().numpy()
        if ref is None:
            ref = chk
        same = bool((chk == ref).all())
        print(json.dumps({"G": g, "GT": t, "split": s, "cfree": combo[3] if len(combo) > 3 else 0, "mode": os.environ["XW_RENDER_MODE"], "two_phase": os.environ["XW_RENDER_TWO_PHASE"], "fill": os.environ["XW_RENDER_SP_FILL"], "ms": round(ms, 4), "GBs": round(n * fb / ms / 1e6, 1), "same_as_first": same}),
              flush=True)
        del sim
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
