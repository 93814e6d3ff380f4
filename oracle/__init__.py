"""CPU oracle for the XWorld2D hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this
package.  The product package `xworld_b200` never does; it only shares the public ABI struct
definitions (xworld_b200._abi mirrors include/xworld_b200.h), which are interface, not behaviour.

`Oracle` wraps oracle/libxw_oracle.so (xw_oracle.c, the C restatement of the reference);
`RefLib` wraps oracle/_ref/libxw_ref.so (the reference's own C++ compiled against header shims).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from xworld_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libxw_oracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libxw_ref.so")

MAXG, MAXD = _abi.XW_MAX_GOALS, _abi.XW_MAX_DIM

(SITE_NAMES, SITE_MAZE, SITE_BLOCKS, SITE_GOAL_LOC, SITE_GOAL_ASSET, SITE_AGENT_LOC, SITE_TASK_A,
 SITE_TASK_B, SITE_TASK_SHUF, SITE_TASK_AGENT) = range(1, 11)
SITE_AGENT_YAW, SITE_GOAL_POSE = 12, 13
YAW_STEPS = 4096


class XoEnv(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("W", C.c_int32),
        ("grid", C.c_uint8 * (MAXD * MAXD)),
        ("agent_x", C.c_int32), ("agent_y", C.c_int32),
        ("agent_yaw", C.c_double),
        ("n_goals", C.c_int32),
        ("goal_x", C.c_int32 * MAXG), ("goal_y", C.c_int32 * MAXG),
        ("goal_icon", C.c_int32 * MAXG), ("goal_name", C.c_int32 * MAXG),
        ("task", C.c_int32), ("stage", C.c_int32), ("event", C.c_int32), ("action_success", C.c_int32),
        ("target_mask", C.c_int32), ("aux0", C.c_int32), ("aux1", C.c_int32), ("aux2", C.c_int32),
        ("steps_in_task", C.c_int32),
        ("num_steps", C.c_int64),
        ("episode", C.c_int32),
        ("n_success", C.c_int32), ("n_failure", C.c_int32), ("success_steps", C.c_int32),
        ("minstd", C.c_uint32),
        ("error", C.c_int32),
        ("env_gid", C.c_int64),
        ("level", C.c_int32), ("dim", C.c_int32), ("check_counter", C.c_int32),
        ("seq_len", C.c_int32 * 5),
        ("seq", (C.c_uint8 * 200) * 5),
        ("goal_yaw", C.c_double * MAXG), ("goal_scale", C.c_double * MAXG), ("goal_offset", C.c_double * MAXG),
    ]


class XoRace(C.Structure):
    _fields_ = [("pos_x", C.c_float), ("pos_y", C.c_float), ("angle", C.c_float), ("steps", C.c_int32), ("minstd", C.c_uint32)]


class XoSimpleGame(C.Structure):
    _fields_ = [("array_size", C.c_int32), ("cur_pos", C.c_int32), ("rewards", C.c_float * 64),
                ("state", C.c_uint8 * 64)]


def build(force=False):
    """Compile xw_oracle.c (and oracle/_ref when /root/reference is present)."""
    srcs = [os.path.join(_HERE, f) for f in ("xw_oracle.c", "xw_oracle_fpv.c", "xw_oracle.h", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "xworld_b200.h"))
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "libxw_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        u32, i32, u64, i64, vp = C.c_uint32, C.c_int32, C.c_uint64, C.c_int64, C.c_void_p
        cfgp, catp, envp = C.POINTER(_abi.XwConfig), C.POINTER(_abi.XwCatalog), C.POINTER(XoEnv)
        L.xo_philox4x32_10.argtypes = [C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        L.xo_draw.argtypes = [u64, i64, u32, u32, u32, u32]
        L.xo_draw.restype = u32
        L.xo_randbelow.argtypes = [u32, u32]
        L.xo_randbelow.restype = u32
        L.xo_std_hash_bytes.argtypes = [C.c_char_p, u64]
        L.xo_std_hash_bytes.restype = u64
        L.xo_minstd_seed_for_thread.argtypes = [i32, i32]
        L.xo_minstd_seed_for_thread.restype = u32
        L.xo_get_rand_ind.argtypes = [C.POINTER(u32), i32]
        L.xo_get_rand_ind.restype = i32
        L.xo_maze.argtypes = [u64, i64, u32, u32, C.c_int, C.c_char_p]
        L.xo_env_init.argtypes = [cfgp, envp, i64]
        L.xo_reset.argtypes = [cfgp, catp, envp]
        L.xo_step.argtypes = [cfgp, catp, envp, i32, i32, C.POINTER(C.c_float), C.POINTER(i32)]
        L.xo_teach.argtypes = [cfgp, catp, envp, i32, C.POINTER(C.c_double)]
        L.xo_resize_tables.argtypes = [C.c_int, C.c_int, vp, vp, vp]
        L.xo_resize_linear_8uc3.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int]
        L.xo_render.argtypes = [cfgp, catp, envp, vp]
        L.xo_frame_dims.argtypes = [cfgp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        dbl = C.c_double
        L.xo_rotation_matrix.argtypes = [C.c_float, C.c_float, dbl, dbl, C.POINTER(dbl)]
        L.xo_warp_affine_8uc3.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.POINTER(dbl), vp]
        L.xo_item_image.argtypes = [vp, dbl, dbl, dbl, vp]
        L.xo_facing_dir.argtypes = [dbl]
        L.xo_image_masking.argtypes = [envp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), vp]
        L.xo_visible_radius.argtypes = [cfgp]
        L.xo_goal_pose.argtypes = [u64, i64, u32, u32, C.c_int, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl)]
        L.xo_render_fpv.argtypes = [cfgp, catp, envp, vp]
        L.xo_batch_reset.argtypes = [cfgp, catp, envp, C.c_int, C.c_int]
        L.xo_batch_step.argtypes = [cfgp, catp, envp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int]
        L.xo_sizeof_env.restype = C.c_int
        L.xo_sg_reset.argtypes = [C.POINTER(XoSimpleGame), C.c_int]
        L.xo_sg_act.argtypes = [C.POINTER(XoSimpleGame), C.c_int]
        L.xo_sg_act.restype = C.c_float
        L.xo_sg_game_over.argtypes = [C.POINTER(XoSimpleGame)]
        L.xo_race_reset.argtypes = [cfgp, C.POINTER(XoRace)]
        L.xo_race_act.argtypes = [cfgp, C.POINTER(XoRace), C.c_int, C.POINTER(C.c_float), C.POINTER(i32)]
        L.xo_race_act.restype = C.c_float
        L.xo_race_take_actions.argtypes = [cfgp, C.POINTER(XoRace), C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(i32)]
        L.xo_race_take_actions.restype = C.c_float
        assert L.xo_sizeof_env() == C.sizeof(XoEnv), (L.xo_sizeof_env(), C.sizeof(XoEnv))
        _lib = L
    return _lib


def draw(seed, env_gid, episode, attempt, site, index):
    return lib().xo_draw(seed, env_gid, episode, attempt, site, index)


def randbelow(u, n):
    return lib().xo_randbelow(u, n)


U8_FIELDS = ("agent_x", "agent_y", "task", "stage", "event", "action_success", "target_mask", "aux0",
             "aux1", "aux2")
I32_FIELDS = ("steps_in_task", "num_steps", "episode", "n_success", "n_failure", "success_steps",
              "minstd")


class Oracle(object):
    """A batch of n CPU environments stepped by the C restatement."""

    def __init__(self, cfg, catalog, n_envs, threads=1):
        self.L = lib()
        self.cfg, self.catalog, self.n, self.threads = cfg, catalog, n_envs, threads
        self.cat_c = catalog.as_c()
        self.envs = (XoEnv * n_envs)()
        for i in range(n_envs):
            self.L.xo_env_init(C.byref(cfg), C.byref(self.envs[i]), cfg.env_id_offset + i)
        oh, ow = C.c_int(), C.c_int()
        self.L.xo_frame_dims(C.byref(cfg), C.byref(oh), C.byref(ow))
        self.out_h, self.out_w = oh.value, ow.value
        self.channels = 1 if cfg.gray else 3

    def reset(self, mask=None):
        if mask is None:
            rc = self.L.xo_batch_reset(C.byref(self.cfg), C.byref(self.cat_c), self.envs, self.n, self.threads)
            assert rc == 0, rc
        else:
            for i in range(self.n):
                if mask[i]:
                    rc = self.L.xo_reset(C.byref(self.cfg), C.byref(self.cat_c), C.byref(self.envs[i]))
                    assert rc == 0, rc

    def step(self, actions, act_rep=1, render=False):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        reward = np.zeros(self.n, np.float32)
        over = np.zeros(self.n, np.int32)
        frames = np.zeros((self.n, self.channels, self.out_h, self.out_w), np.uint8) if render else None
        rc = self.L.xo_batch_step(
            C.byref(self.cfg), C.byref(self.cat_c), self.envs, self.n, actions.ctypes.data, act_rep,
            reward.ctypes.data, over.ctypes.data, frames.ctypes.data if render else None, self.threads)
        assert rc == 0, rc
        return reward, over, frames

    def render(self, idx=None):
        idx = range(self.n) if idx is None else idx
        out = np.zeros((len(idx), self.channels, self.out_h, self.out_w), np.uint8)
        for k, i in enumerate(idx):
            self.L.xo_render(C.byref(self.cfg), C.byref(self.cat_c), C.byref(self.envs[i]), out[k].ctypes.data)
        return out

    def field(self, name):
        """State in the layout xw_get_field uses."""
        n, H, W = self.n, self.cfg.height, self.cfg.width
        if name == "grid":
            return np.array([list(e.grid)[:H * W] for e in self.envs], np.uint8)
        if name in ("goal_x", "goal_y"):
            return np.array([list(getattr(e, name)) for e in self.envs], np.uint8)
        if name == "goal_icon":
            return np.array([list(e.goal_icon) for e in self.envs], np.int32)
        if name == "facing":  # XItem::get_item_facing_dir of the agent's yaw: 0 right, 1 down, 2 left, 3 up
            return np.array([self.L.xo_facing_dir(e.agent_yaw) for e in self.envs], np.uint8)
        if name in ("goal_scale", "goal_offset", "goal_yaw"):
            return np.array([list(getattr(e, name)) for e in self.envs], np.float64)
        if name == "goal_yaw_idx":  # yaw = 4 * PI_2 * idx / 4096 (xo_goal_pose); defaults (fully observed) read 1024
            return np.array([[int(round(y / (1.5707963 * 4) * YAW_STEPS)) for y in e.goal_yaw] for e in self.envs], np.uint16)
        if name == "level":
            return np.array([e.level for e in self.envs], np.uint8)
        if name == "check_counter":
            return np.array([e.check_counter for e in self.envs], np.int32)
        if name == "win_len":
            return np.array([list(e.seq_len) for e in self.envs], np.uint8)
        if name == "win_sum":
            return np.array([[sum(e.seq[t][:e.seq_len[t]]) for t in range(5)] for e in self.envs], np.uint8)
        if name in U8_FIELDS:
            return np.array([getattr(e, name) for e in self.envs], np.uint8)
        if name in I32_FIELDS:
            return np.array([getattr(e, name) for e in self.envs], np.int64).astype(np.int32)
        raise KeyError(name)


class RefLib(object):
    """The reference's own C++ (compiled unmodified against oracle/shim) -- see ref_driver.cpp."""

    def __init__(self):
        if not os.path.exists(REF_LIB):
            raise RuntimeError("oracle/_ref/libxw_ref.so missing (built from /root/reference by `make -C oracle ref`)")
        self.L = C.CDLL(REF_LIB)
