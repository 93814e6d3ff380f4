"""Parity harness shared by the CPU (hostsim) and GPU (C ABI) tests: drives a backend and the
oracle with the same seeded action stream and compares every output bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle
from xworld_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))

U8 = ["agent_x", "agent_y", "facing", "task", "stage", "event", "action_success", "target_mask", "aux0", "aux1", "aux2"]
I32 = ["steps_in_task", "num_steps", "episode", "n_success", "n_failure", "success_steps", "minstd"]

CONFIGS = {
    # BASELINE.json configs: C2, C3, C4 + the reference's own 8x8 default
    "c2_nav3d_7x7_84": dict(height=7, width=7, n_goals=4, n_blocks=12, rules=_abi.XW_RULES_NAV3D),
    "c3_nav2d_11x11_84": dict(height=11, width=11, n_goals=4, n_blocks=30, rules=_abi.XW_RULES_NAV2D, out_h=84, out_w=84,
                              max_steps=242),
    "c4_nav3d_15x15_128": dict(height=15, width=15, n_goals=4, n_blocks=56, rules=_abi.XW_RULES_NAV3D, out_h=128, out_w=128),
    "ref_nav3d_8x8_96": dict(height=8, width=8, n_goals=4, n_blocks=16, rules=_abi.XW_RULES_NAV3D),
    # XWorldNav's level schedule (the check period is shortened so that levels change within a test)
    "curriculum_nav3d_8x8_96": dict(height=8, width=8, n_goals=4, n_blocks=16, rules=_abi.XW_RULES_NAV3D, curriculum=0.1,
                                    curriculum_check_period=5),
    "curriculum_nav2d_8x8_96": dict(height=8, width=8, n_goals=4, n_blocks=16, rules=_abi.XW_RULES_NAV2D, curriculum=0.1,
                                    start_level=2, max_steps=40),
    "ref_nav2d_8x8_96": dict(height=8, width=8, n_goals=4, n_blocks=16, rules=_abi.XW_RULES_NAV2D, max_steps=64),
    # the first-person view.  11x11 with visible_radius 7 is the reference-native 84x84 frame of BASELINE config 3's map
    "fpv_nav2d_11x11_vr7_84": dict(height=11, width=11, n_goals=4, n_blocks=30, rules=_abi.XW_RULES_NAV2D, visible_radius=7, max_steps=242),
    "fpv_nav3d_8x8_vr3_84": dict(height=8, width=8, n_goals=4, n_blocks=16, rules=_abi.XW_RULES_NAV3D, visible_radius=3),
    "fpv_nav3d_7x7_vr7_84": dict(height=7, width=7, n_goals=4, n_blocks=12, rules=_abi.XW_RULES_NAV3D, visible_radius=7),
    "fpv_nav3d_8x8_vr5_80": dict(height=8, width=8, n_goals=4, n_blocks=16, rules=_abi.XW_RULES_NAV3D, visible_radius=5),
    "fpv_nav3d_15x15_vr9_81": dict(height=15, width=15, n_goals=4, n_blocks=56, rules=_abi.XW_RULES_NAV3D, visible_radius=9),
    "fpv_nav3d_7x7_vr1_84": dict(height=7, width=7, n_goals=4, n_blocks=12, rules=_abi.XW_RULES_NAV3D, visible_radius=1),
    # --task_mode=one_channel (the reference's Python default, py_simulator.cpp:128-130)
    "one_channel_nav3d_8x8_96": dict(height=8, width=8, n_goals=4, n_blocks=16, rules=_abi.XW_RULES_NAV3D, max_steps=60,
                                     task_mode=_abi.XW_TASK_ONE_CHANNEL),
    "one_channel_nav2d_8x8_96": dict(height=8, width=8, n_goals=4, n_blocks=16, rules=_abi.XW_RULES_NAV2D, max_steps=90,
                                     task_mode=_abi.XW_TASK_ONE_CHANNEL),
}


def make_cfg(name, **over):
    kw = dict(CONFIGS[name])
    kw.setdefault("seed", 1234)
    kw.setdefault("simulator_seed", 1)
    kw.update(over)
    return _abi.default_config(**kw)


def frame_dims(cfg):
    """XWorldSimulator::init (xworld_simulator.cpp:48-68)."""
    h, w = cfg.height * 12, cfg.width * 12
    if cfg.visible_radius > 0:
        vr = min(cfg.visible_radius, cfg.height)
        h = w = vr * (84 // vr)
    return cfg.out_h or h, cfg.out_w or w


def n_actions_of(cfg):
    return 6 if cfg.visible_radius > 0 else 4


def actions_for(step, n, n_actions, seed=99):
    """i.i.d. uniform actions, a pure function of (seed, step)."""
    return np.random.RandomState((seed * 1000003 + step) % (2 ** 31)).randint(0, n_actions, size=n).astype(np.int32)


class HostSim(object):
    """tests/hostsim/libhostsim.so: the engine's per-env device functions compiled for the host."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            src = os.path.join(HERE, "hostsim", "hostsim.cpp")
            so = os.path.join(HERE, "hostsim", "libhostsim.so")
            deps = [src] + [os.path.join(HERE, "..", "xworld_b200", "csrc", f)
                            for f in os.listdir(os.path.join(HERE, "..", "xworld_b200", "csrc"))]
            if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
                subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src])
            L = C.CDLL(so)
            L.hs_create.restype = C.c_void_p
            L.hs_create.argtypes = [C.POINTER(_abi.XwConfig), C.POINTER(_abi.XwCatalog), C.c_int]
            L.hs_destroy.argtypes = [C.c_void_p]
            L.hs_reset.argtypes = [C.c_void_p, C.c_void_p]
            L.hs_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
            L.hs_render.argtypes = [C.c_void_p, C.c_void_p]
            L.hs_render_mode.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            L.hs_sp_ok.argtypes = [C.c_void_p]
            L.hs_get_field.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
            L.hs_build_phase_atlas.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            L.hs_fast_ok.argtypes = [C.c_void_p]
            L.hs_threads.argtypes = [C.c_void_p]
            L.hs_set_fpv_env.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, cfg, catalog, n):
        self.L = self.lib()
        self.cfg, self.catalog, self.n = cfg, catalog, n
        self.cat_c = catalog.as_c() if catalog is not None else None
        self.h = self.L.hs_create(C.byref(cfg), C.byref(self.cat_c) if catalog is not None else None, n)
        self.out_h, self.out_w = frame_dims(cfg)
        self._atlas_icons = set()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hs_destroy(self.h)
            self.h = None

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self.L.hs_reset(self.h, None if m is None else m.ctypes.data)

    def step(self, actions, act_rep=1, render=False):
        a = np.ascontiguousarray(actions, np.int32)
        r = np.zeros(self.n, np.float32)
        o = np.zeros(self.n, np.int32)
        self.L.hs_step(self.h, a.ctypes.data, act_rep, r.ctypes.data, o.ctypes.data)
        return r, o, (self.render() if render else None)

    def render(self, mode=None):
        """mode None: what the engine would pick (sparse painter when the geometry allows); 0: the plan
        compositor; 1: the sparse painter."""
        icons = set(self.field("goal_icon")[:, :self.cfg.n_goals].ravel().tolist())
        icons |= {self.catalog.brick_icon, self.catalog.agent_icon}
        if self.cfg.visible_radius > 0:
            icons = set()  # the first-person view has no phase atlas
        if not icons <= self._atlas_icons:
            self._atlas_icons |= icons
            arr = np.array(sorted(self._atlas_icons), np.int32)
            self.L.hs_build_phase_atlas(self.h, arr.ctypes.data, len(arr))
        out = np.zeros((self.n, 3, self.out_h, self.out_w), np.uint8)
        if mode is None:
            self.L.hs_render(self.h, out.ctypes.data)
        else:
            self.L.hs_render_mode(self.h, out.ctypes.data, mode)
        return out

    def field(self, name):
        n = self.n
        if name == "grid":
            out = np.zeros((n, self.cfg.height * self.cfg.width), np.uint8)
        elif name in ("goal_x", "goal_y"):
            out = np.zeros((n, _abi.XW_MAX_GOALS), np.uint8)
        elif name in ("goal_icon", "goal_name"):
            out = np.zeros((n, _abi.XW_MAX_GOALS), np.int32)
        elif name == "goal_yaw":
            out = np.zeros((n, _abi.XW_MAX_GOALS), np.uint16)
        elif name in ("goal_scale", "goal_offset"):
            out = np.zeros((n, _abi.XW_MAX_GOALS), np.float64)
        elif name in ("win_len", "win_sum"):
            out = np.zeros((n, 5), np.uint8)
        elif name in U8 or name == "level":
            out = np.zeros(n, np.uint8)
        elif name == "state":
            out = np.zeros((n, 4), np.float32)
        elif name in ("pos_x", "pos_y", "angle"):
            out = np.zeros(n, np.float32)
        else:
            out = np.zeros(n, np.int32)
        rc = self.L.hs_get_field(self.h, name.encode(), out.ctypes.data)
        assert rc == 0, name
        return out


def compare_state(backend, orc, tag=""):
    G = orc.cfg.n_goals
    for f in ["grid"] + U8 + I32:
        a, b = backend.field(f), orc.field(f)
        if f == "minstd":
            a, b = a.astype(np.uint32), b.astype(np.uint32)
        bad = np.nonzero(np.atleast_2d(a.reshape(len(a), -1) != b.reshape(len(b), -1)).any(axis=1))[0]
        assert bad.size == 0, "%s field %s differs for envs %s: got %s want %s" % (
            tag, f, bad[:5], a[bad[:3]], b[bad[:3]])
    for f in ["goal_x", "goal_y", "goal_icon"]:
        a, b = backend.field(f)[:, :G], orc.field(f)[:, :G]
        assert (a == b).all(), "%s field %s differs" % (tag, f)
    if orc.cfg.visible_radius > 0:  # the goals' poses: the yaw's grid index, scale and offset as bit-equal doubles
        assert (backend.field("goal_yaw")[:, :G] == orc.field("goal_yaw_idx")[:, :G]).all(), "%s goal_yaw differs" % tag
        for f in ["goal_scale", "goal_offset"]:
            a, b = backend.field(f)[:, :G], orc.field(f)[:, :G]
            assert (a.view(np.uint64) == b.view(np.uint64)).all(), "%s field %s differs" % (tag, f)
    if orc.cfg.curriculum != 0:
        for f in ["level", "check_counter", "win_len", "win_sum"]:
            a, b = backend.field(f), orc.field(f)
            assert (a == b).all(), "%s curriculum field %s differs: got %s want %s" % (tag, f, a[:4], b[:4])


def run_parity(backend, orc, n_steps, render_every=0, act_rep=1, check_state_every=1, auto_reset=False, seed=99):
    """Same actions into both; compare reward bits, game_over, state, frames.  Returns stats."""
    n = orc.n
    backend.reset()
    orc.reset()
    compare_state(backend, orc, "after reset")
    stats = {"events": {}, "frames": 0, "resets": 0}
    if render_every:
        fa, fb = backend.render(), orc.render()
        assert (fa == fb).all(), "first frame differs: %d px" % (fa != fb).sum()
        stats["frames"] += n
    for s in range(n_steps):
        a = actions_for(s, n, n_actions_of(orc.cfg), seed)
        do_render = bool(render_every) and (s % render_every == 0)
        r1, o1, f1 = backend.step(a, act_rep, render=do_render)
        r2, o2, f2 = orc.step(a, act_rep, render=do_render)
        assert (r1.view(np.uint32) == r2.view(np.uint32)).all(), "step %d reward bits differ: %s vs %s" % (
            s, r1[(r1.view(np.uint32) != r2.view(np.uint32))][:4], r2[(r1.view(np.uint32) != r2.view(np.uint32))][:4])
        assert (o1 == o2).all(), "step %d game_over differs" % s
        for v in np.unique(o2):
            stats["events"][int(v)] = stats["events"].get(int(v), 0) + int((o2 == v).sum())
        if not auto_reset and (o2 != 0).any():  # reference behaviour: the caller resets finished games
            m = (o2 != 0)
            backend.reset(m)
            orc.reset(m)
            stats["resets"] += int(m.sum())
            if do_render:
                f1, f2 = backend.render(), orc.render()
        if check_state_every and s % check_state_every == 0:
            compare_state(backend, orc, "step %d" % s)
        if do_render:
            assert (f1 == f2).all(), "step %d: %d frame bytes differ" % (s, (f1 != f2).sum())
            stats["frames"] += n
    compare_state(backend, orc, "final")
    return stats


def run_partial_parity(backend, orc, n_steps, p_step=0.6, render_every=0, seed=5):
    """XW_ACTION_NONE: every step a random subset of the envs acts, the others sit it out.  The oracle steps only
    the subset (the reference: an env that is not stepped is simply not called); state of ALL envs, and the reward /
    game_over of the subset, must agree; finished games of the subset are reset by the caller."""
    import ctypes as C
    n = orc.n
    backend.reset()
    orc.reset()
    rng = np.random.RandomState(seed)
    stepped = 0
    for s in range(n_steps):
        m = rng.rand(n) < p_step
        a = actions_for(s, n, n_actions_of(orc.cfg), seed)
        a_eng = np.where(m, a, _abi.XW_ACTION_NONE).astype(np.int32)
        r1, o1, f1 = backend.step(a_eng, 1, render=bool(render_every) and s % render_every == 0)
        done = np.zeros(n, np.uint8)
        for i in np.nonzero(m)[0]:
            r, ov = C.c_float(), C.c_int32()
            rc = orc.L.xo_step(C.byref(orc.cfg), C.byref(orc.cat_c), C.byref(orc.envs[i]), int(a[i]), 1, C.byref(r), C.byref(ov))
            assert rc == 0
            assert np.float32(r.value).view(np.uint32) == r1[i].view(np.uint32) and ov.value == o1[i], (s, i)
            done[i] = ov.value != 0
        stepped += int(m.sum())
        if f1 is not None:
            f2 = orc.render()
            assert (f1 == f2).all(), "step %d: frames differ" % s
        if done.any():
            backend.reset(done)
            orc.reset(done)
        compare_state(backend, orc, "partial step %d" % s)
    return stepped
