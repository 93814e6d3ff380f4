"""Adapter giving the real engine (through the C ABI, via xworld_b200.Simulator) the interface
tests/parity.py drives."""
import ctypes as C

import numpy as np

from xworld_b200 import _abi
from xworld_b200.simulator import Simulator


class EngineBackend(object):
    def __init__(self, cfg, catalog, n, device=-1):
        self.cfg, self.n = cfg, n
        self.sim = Simulator("xworld" if cfg.game == _abi.XW_GAME_XWORLD else "simple_race", cfg, catalog, n, device)
        import torch
        self.torch = torch

    def reset(self, mask=None):
        self.sim.reset_game(mask)

    def step(self, actions, act_rep=1, render=False):
        t = self.torch
        a = t.from_numpy(np.ascontiguousarray(actions, np.int32)).cuda()
        sim = self.sim
        with t.cuda.device(sim._dev):
            rc = sim._lib.xw_step(sim._h, a.data_ptr(), act_rep, sim._d_reward.data_ptr(), sim._d_over.data_ptr(),
                                  sim._screen.data_ptr() if render else None, sim._stream())
            assert rc == 0, sim._lib.xw_last_error()
        r = sim._d_reward.cpu().numpy()
        o = sim._d_over.cpu().numpy()
        return r, o, (sim._screen.cpu().numpy() if render else None)

    def render(self):
        sim = self.sim
        with self.torch.cuda.device(sim._dev):
            rc = sim._lib.xw_render(sim._h, sim._screen.data_ptr(), sim._stream())
            assert rc == 0, sim._lib.xw_last_error()
        return sim._screen.cpu().numpy()

    def field(self, name):
        return self.sim.get_field(name)
