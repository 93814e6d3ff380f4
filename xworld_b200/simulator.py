"""Host-side mirror of the reference's Python API, over the C ABI.

    from xworld_b200 import Simulator
    sim = Simulator.create("xworld", {"xwd_conf_path": ".../confs/navigation2d.json", "n_envs": 65536})

keeps the method names and semantics of `py_simulator.Simulator` (python/py_simulator.cpp:310-329):
create, reset_game, game_over, get_num_actions, get_lives, get_screen_out_dimensions, take_actions,
take_action, get_state, get_num_steps.  With n_envs == 1 (the default) the return types are the
reference's (float reward, "alive|dead|..." string, dict with a [0,1] float screen list); with
n_envs > 1 the same calls take / return arrays, and CUDA tensors stay on the device.

There is no CPU fallback for xworld / simple_race: without the CUDA library or a GPU, create() raises.
"""
import ctypes as C
import json
import os

import numpy as np

from . import _abi
from .catalog import Catalog

_CODE_NAMES = ((_abi.XW_MAX_STEP, "max_step"), (_abi.XW_DEAD, "dead"), (_abi.XW_SUCCESS, "success"),
               (_abi.XW_LOST_LIFE, "lost_life"))


def decode_game_over_code(code):
    """GameSimulator::decode_game_over_code (simulator.cpp:125-144)."""
    if code == 0:
        return "alive"
    return "|".join(n for bit, n in _CODE_NAMES if code & bit)


def _opt(opts, key, required, default):
    """extract_py_dict_val (py_simulator.cpp:36-56): missing required key -> RuntimeError."""
    if key in opts:
        return opts[key]
    if required:
        raise RuntimeError("Key '%s' is required" % key)
    return default


def rules_from_conf(conf):
    """Which teacher rule set a conf json selects (teacher.cpp:70-99 reads task_groups)."""
    groups = conf.get("task_groups", {})
    tasks = [t for g in groups.values() for t in g.get("tasks", {})]
    if any(t.startswith("XWorld3DNav") for t in tasks):
        return _abi.XW_RULES_NAV3D
    if any(t.startswith("XWorldNav") for t in tasks):
        return _abi.XW_RULES_NAV2D
    raise RuntimeError("conf has no navigation task group this engine implements: %s" % sorted(groups))


SENTENCE_FIELDS = ("task", "stage", "event", "aux0", "aux1", "goal_name", "goal_icon", "episode", "steps_in_task", "num_steps")


def sentence_for_state(lib, cfg, catalog, env_id, st):
    """The teacher's sentence of one env from its state fields (SENTENCE_FIELDS): which sentence the last teach()
    produced (see Simulator.sentences) and its slots -- goal names, direction, colour -- then xw_sentence_compose."""
    q = _abi.XwSentenceQuery(rules=cfg.rules, task=int(st["task"]), kind=_abi.XW_SENT_START, direction=0, name1=None, name2=None,
                             color=None, seed=cfg.seed, env_id=env_id, episode=int(st["episode"]) & 0xffffffff, salt=0)
    stage, event, a0 = int(st["stage"]), int(st["event"]), int(st["aux0"])
    name = lambda g: catalog.names[int(st["goal_name"][g])].encode()
    if cfg.rules == _abi.XW_RULES_NAV3D:
        if stage == _abi.XW_STAGE_IDLE or (stage == _abi.XW_STAGE_TERMINAL and event == _abi.XW_EVENT_NONE):
            return ""
        if stage == _abi.XW_STAGE_TERMINAL:  # the step that ended the episode (xworld3d_task.py:456-482)
            q.kind = {_abi.XW_EVENT_CORRECT_GOAL: _abi.XW_SENT_CORRECT, _abi.XW_EVENT_WRONG_GOAL: _abi.XW_SENT_WRONG,
                      _abi.XW_EVENT_TIME_UP: _abi.XW_SENT_TIMEUP}[event]
        elif q.task == 2:    # XW_T3_BETWEEN: G1, G2
            q.name1, q.name2 = name(a0 & 15), name(a0 >> 4)
        else:                # Target / Near / Direction / Avoid: G (Direction: the referent, P = "aux1")
            q.name1 = name(a0)
            q.direction = int(st["aux1"]) if q.task == 3 else 0
    else:
        if event == _abi.XW_EVENT_CORRECT_GOAL:
            q.kind = _abi.XW_SENT_CORRECT
        elif stage == _abi.XW_STAGE_NAVIGATION and int(st["steps_in_task"]) == 0:
            q.name1 = name(a0)
            q.color = catalog.icon_meta[int(st["goal_icon"][a0])]["color"].encode()
            q.salt = int(st["num_steps"]) & 0x3fff
        else:
            return ""
    buf = C.create_string_buffer(256)
    rc = lib.xw_sentence_compose(C.byref(q), buf, len(buf))
    if rc < 0:
        raise RuntimeError("xworld_b200 error %d: %s" % (rc, lib.xw_last_error().decode()))
    return buf.value.decode()


class Simulator(object):
    def __init__(self, name, cfg, catalog, n_envs, device):
        self._lib = _abi.load()
        self.name, self.cfg, self.catalog, self.n_envs = name, cfg, catalog, n_envs
        h = C.c_void_p()
        rc = self._lib.xw_create(C.byref(cfg), C.byref(catalog.as_c()) if catalog is not None else None,
                                 n_envs, device, C.byref(h))
        if rc != 0:
            raise RuntimeError("xw_create failed (%d): %s" % (rc, self._lib.xw_last_error().decode()))
        self._h = h
        dims = [C.c_int32() for _ in range(4)]
        self._lib.xw_screen_dims(h, *[C.byref(d) for d in dims])
        self.height, self.width, self.channels, self.context = [d.value for d in dims]
        self._on_gpu = cfg.game != _abi.XW_GAME_SIMPLE_GAME
        self._last_over = np.zeros(n_envs, np.int32)
        self._last_reward = np.zeros(n_envs, np.float32)
        self._stale = False
        self._screen = None  # device tensor (gpu games) or numpy (simple_game)
        self._torch = None
        if self._on_gpu:
            import torch
            self._torch = torch
            self._dev = torch.device("cuda", torch.cuda.current_device() if device < 0 else device)
            dt = torch.float32 if cfg.game == _abi.XW_GAME_SIMPLE_RACE else torch.uint8
            self._screen = torch.zeros((n_envs, self.context * self.channels, self.height, self.width), dtype=dt,
                                       device=self._dev)
            self._d_reward = torch.zeros(n_envs, dtype=torch.float32, device=self._dev)
            self._d_over = torch.zeros(n_envs, dtype=torch.int32, device=self._dev)
        else:
            self._screen = np.zeros((n_envs, self.width), np.uint8)

    # ------------------------------------------------------------------ factory
    @staticmethod
    def create(name, opts=None):
        """py_simulator.Simulator.create(name, dict) (py_simulator.cpp:169-191)."""
        opts = dict(opts or {})
        n_envs = int(opts.pop("n_envs", 1))
        device = int(opts.pop("device", -1))
        catalog = None
        if name == "simple_game":
            cfg = _abi.default_config(game=_abi.XW_GAME_SIMPLE_GAME, array_size=int(_opt(opts, "array_size", True, 0)))
        elif name == "simple_race":
            cfg = _abi.default_config(
                game=_abi.XW_GAME_SIMPLE_RACE,
                track_type={"straight": 0, "circle": 1}[_opt(opts, "track_type", False, "straight")],
                track_width=float(_opt(opts, "track_width", True, 20.0)),
                track_length=float(_opt(opts, "track_length", True, 100.0)),
                track_radius=float(_opt(opts, "track_radius", True, 30.0)),
                race_full_manouver=int(bool(_opt(opts, "race_full_manouver", False, False))),
                race_random=int(bool(_opt(opts, "random", False, False))),
                difficulty={"easy": 0, "hard": 1}[_opt(opts, "difficulty", False, "easy")],
                reward_scale=float(opts.get("reward_scale", 1.0)))
        elif name == "xworld":
            conf_path = _opt(opts, "xwd_conf_path", True, "")
            with open(conf_path) as f:
                conf = json.load(f)
            if conf.get("map") != "XWorldNav":
                raise RuntimeError("map '%s' is not implemented (XWorldNav only)" % conf.get("map"))
            task_mode = _opt(opts, "task_mode", False, "one_channel")  # py default, py_simulator.cpp:128-130
            if task_mode not in ("lang_acquisition", "one_channel"):
                raise RuntimeError("task_mode '%s' is not implemented (lang_acquisition | one_channel)" % task_mode)
            color = bool(_opt(opts, "color", False, False))            # py default, py_simulator.cpp:135
            curriculum = float(_opt(opts, "curriculum", False, 0))  # py_simulator.cpp:127
            catalog = opts.get("catalog")
            if catalog is None:
                item_path = opts.get("item_path")
                if item_path is None:  # xworld.cpp:86: <games/xworld>/<item_path>
                    for rel in ("..", os.path.join("..", "games", "xworld")):
                        cand = os.path.join(os.path.dirname(os.path.abspath(conf_path)), rel, conf["item_path"])
                        if os.path.isdir(cand):
                            item_path = cand
                            break
                if item_path is None:
                    raise RuntimeError("cannot locate item_path '%s' next to %s; pass opts['item_path'] or "
                                       "opts['catalog']" % (conf["item_path"], conf_path))
                catalog = Catalog.from_item_path(item_path)
            dim = int(opts.get("map_size", 8))  # XWorldNav.py:10-11 hard-codes 8
            default_blocks = {7: 12, 8: 16, 11: 30, 15: 56}.get(dim, 2 * dim)
            cfg = _abi.default_config(
                game=_abi.XW_GAME_XWORLD, height=dim, width=dim,
                n_goals=int(opts.get("n_goals", 4)), n_blocks=int(opts.get("n_blocks", default_blocks)),
                rules=rules_from_conf(conf), out_h=int(opts.get("out_h", 0)), out_w=int(opts.get("out_w", 0)),
                context=int(_opt(opts, "context", False, 1)), visible_radius=int(_opt(opts, "visible_radius", False, 0)),
                max_steps=int(opts.get("max_steps", 0)), max_steps_factor=int(opts.get("max_steps_factor", 10)),
                curriculum=curriculum, curriculum_check_period=int(opts.get("curriculum_check_period", 0)),
                start_level=int(opts.get("start_level", 0)),
                task_mode=_abi.XW_TASK_ONE_CHANNEL if task_mode == "one_channel" else _abi.XW_TASK_LANG_ACQUISITION,
                gray=0 if color else 1)
        else:
            raise RuntimeError("Unrecognized game type: " + name)
        cfg.auto_reset = int(bool(opts.get("auto_reset", False)))
        cfg.seed = int(opts.get("seed", 0))
        cfg.simulator_seed = int(opts.get("simulator_seed", 0))
        cfg.env_id_offset = int(opts.get("env_id_offset", 0))
        if "max_steps" in opts:
            cfg.max_steps = int(opts["max_steps"])
        return Simulator(name, cfg, catalog, n_envs, device)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.xw_destroy(h)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("xworld_b200 error %d: %s" % (rc, self._lib.xw_last_error().decode()))

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self._dev).cuda_stream)

    # ------------------------------------------------------------------ reference API
    def reset_game(self, mask=None):
        """SimulatorInterface::reset_game (simulator_interface.cpp:95-105)."""
        if not self._on_gpu:
            m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
            self._check(self._lib.xw_reset_host(self._h, None if m is None else m.ctypes.data, self._screen.ctypes.data))
        else:
            t = self._torch
            dm = None
            if mask is not None:
                dm = t.as_tensor(mask).to(device=self._dev, dtype=t.uint8).contiguous()
            with t.cuda.device(self._dev):
                if self.context > 1 and mask is None:
                    self._screen.zero_()
                self._check(self._lib.xw_reset(self._h, None if dm is None else dm.data_ptr(), self._stream()))
                if self.cfg.game == _abi.XW_GAME_XWORLD:
                    self._check(self._lib.xw_render(self._h, self._screen.data_ptr(), self._stream()))
                if dm is None:
                    self._d_over.zero_()
                    self._d_reward.zero_()
                else:  # only the envs that were reset are "alive" again
                    self._d_over.masked_fill_(dm.bool(), 0)
                    self._d_reward.masked_fill_(dm.bool(), 0)
        if mask is None:
            self._last_over[:] = 0
        else:
            self._last_over[np.asarray(mask).astype(bool)] = 0

    def _refresh(self):
        """After device-side steps (take_actions with a CUDA tensor) the host copies of reward / game_over are stale."""
        if self._stale:
            self._last_over = self._d_over.cpu().numpy()
            self._last_reward = self._d_reward.cpu().numpy()
            self._stale = False

    def game_over(self):
        """'alive' | 'max_step' | 'dead' | 'success' | 'lost_life' (|-joined); a list when n_envs > 1."""
        self._refresh()
        if self.n_envs == 1:
            return decode_game_over_code(int(self._last_over[0]))
        return [decode_game_over_code(int(c)) for c in self._last_over]

    def game_over_codes(self):
        self._refresh()
        return self._last_over.copy()

    def get_num_actions(self):
        return int(self._lib.xw_num_actions(self._h))

    def get_lives(self):
        """XWorldSimulator::get_lives: game_over() ? 0 : 1 (xworld_simulator.cpp:506)."""
        self._refresh()
        lives = (self._last_over == 0).astype(np.int32)
        return int(lives[0]) if self.n_envs == 1 else lives

    def get_screen_out_dimensions(self):
        return [self.height, self.width, self.channels, self.context]

    def get_num_steps(self):
        out = np.zeros(self.n_envs, np.int64)
        self._check(self._lib.xw_num_steps(self._h, out.ctypes.data))
        return int(out[0]) if self.n_envs == 1 else out

    def take_actions(self, actions, act_rep=1, show_screen=False):
        """SimulatorInterface::take_actions (simulator_interface.cpp:126-137).

        actions: the reference's dict {"action": int[, "pred_sentence": str]} (n_envs == 1), a host
        int array [n_envs], or a CUDA int32 tensor [n_envs] (stays on the device: returns CUDA tensors
        (reward, game_over) and refreshes the device screen)."""
        if show_screen:
            raise RuntimeError("show_screen is not supported (no GUI in the batched engine)")
        t = self._torch
        if t is not None and isinstance(actions, t.Tensor) and actions.is_cuda:
            a = actions.to(dtype=t.int32).contiguous()
            if a.numel() != self.n_envs:
                raise RuntimeError("expected %d actions" % self.n_envs)
            with t.cuda.device(self._dev):
                self._check(self._lib.xw_step(self._h, a.data_ptr(), int(act_rep), self._d_reward.data_ptr(),
                                              self._d_over.data_ptr(), self._screen.data_ptr(), self._stream()))
            self._stale = True
            return self._d_reward, self._d_over
        if isinstance(actions, dict):
            if len(actions) == 0:
                raise RuntimeError("You can't take an empty action")
            a = np.full(self.n_envs, int(actions.get("action", 0)), np.int32)
        else:
            a = np.ascontiguousarray(actions, np.int32).reshape(-1)
            if a.size != self.n_envs:
                raise RuntimeError("expected %d actions" % self.n_envs)
        self._refresh()
        r = np.array(self._last_reward, np.float32)  # envs given XW_ACTION_NONE keep their last reward / game_over
        o = np.array(self._last_over, np.int32)
        if not self._on_gpu:
            self._check(self._lib.xw_step_host(self._h, a.ctypes.data, int(act_rep), r.ctypes.data, o.ctypes.data,
                                               self._screen.ctypes.data))
        else:
            da = t.from_numpy(a).to(self._dev)
            with t.cuda.device(self._dev):
                self._check(self._lib.xw_step(self._h, da.data_ptr(), int(act_rep), self._d_reward.data_ptr(),
                                              self._d_over.data_ptr(), self._screen.data_ptr(), self._stream()))
            stepped = a != _abi.XW_ACTION_NONE   # the kernel leaves the slots of the others untouched: keep the host's values
            r = np.where(stepped, self._d_reward.cpu().numpy(), r)
            o = np.where(stepped, self._d_over.cpu().numpy(), o)
        self._last_reward, self._last_over = r, o
        return float(r[0]) if self.n_envs == 1 else r

    def take_action(self, actions, show_screen=False):
        return self.take_actions(actions, 1, show_screen)

    def screen(self):
        """The raw screen bytes: CUDA uint8 tensor [n_envs, context*C, H, W] (planes B, G, R)."""
        return self._screen

    def get_state(self):
        """PySimulatorInterface::get_state (py_simulator.cpp:246-285).  n_envs == 1: the reference dict
        (uint8 screen scaled by 1/255 into a float list); n_envs > 1: {"screen": raw tensor}."""
        if self.n_envs > 1:
            return {"screen": self._screen}
        if self._on_gpu:
            scr = self._screen[0].reshape(-1).cpu().numpy()
        else:
            scr = self._screen[0]
        d = {}
        if self.cfg.game == _abi.XW_GAME_SIMPLE_RACE:
            d["screen"] = [float(x) for x in scr]
        else:
            scale = np.float32(1 / 255.0)
            d["screen"] = [float(np.float32(x) * scale) for x in scr]
        if self.cfg.game == _abi.XW_GAME_XWORLD:
            d["sentence"] = self.sentences([0])[0]  # the teacher's sentence of the last teach()
            # parse_extra_sim_info (py_simulator.cpp:222-244): "<pid>|k:v,k:v,..." -> the k:v pairs, as strings
            for kv in self.get_extra_info(0).split("|", 1)[1].split(","):
                k, v = kv.split(":", 1)
                d[k] = v
        return d

    def get_extra_info(self, env=0):
        """XWorldSimulator::get_extra_info (xworld_simulator.cpp:495-504): "<id>|task:..,event:..,height:..,width:..";
        the env's global id stands where the reference prints its process id."""
        buf = C.create_string_buffer(256)
        rc = self._lib.xw_extra_info(self._h, int(env), buf, len(buf))
        if rc < 0:
            self._check(rc)
        return buf.value.decode()

    def get_world_dimensions(self):
        """SimulatorInterface::get_world_dimensions (simulator_interface.cpp:163-167)."""
        X, Y, Z = C.c_double(0), C.c_double(0), C.c_double(0)
        self._check(self._lib.xw_world_dimensions(self._h, C.byref(X), C.byref(Y), C.byref(Z)))
        return X.value, Y.value, Z.value

    def teacher_report_task_performance(self):
        """Teacher::report_task_performance (teacher.cpp:175-200): the lines the reference logs, returned (and the raw
        counters): "=== <task> ===" / "=== <S>(S)/<F>(F) -> <rate>@<steps per success>" for every task class that occurred."""
        i64 = C.c_int64 * 8
        s_, f_, st_ = i64(), i64(), i64()
        names = (C.c_char_p * 8)()
        nt = self._lib.xw_task_performance(self._h, s_, f_, st_, 8, names)
        if nt < 0:
            self._check(nt)
        lines, raw = [], {}
        for t in range(nt):
            name = names[t].decode()
            raw[name] = (int(s_[t]), int(f_[t]), int(st_[t]))
            lines.append("=== %s ===" % name)
            if s_[t] + f_[t] == 0:
                continue
            per = float(st_[t]) / s_[t] if s_[t] > 0 else -1
            lines.append("=== %d(S)/%d(F) -> %g@%g" % (s_[t], f_[t], float(s_[t]) / (s_[t] + f_[t]), per))
        return lines, raw

    # ------------------------------------------------------------------ teacher language channel
    def sentences(self, envs=None):
        """The sentence the teacher's last teach() call produced, per env (XWorldSimulator::define_state_specs,
        xworld_simulator.cpp:486-493; CFG.generate over the task grammars, see include/xworld_b200.h):
          navigation2d.json tasks: the command while the episode runs (XWorld3DNavTarget*.py return self.sentence
            from idle() and from every navigation_reward()), "Well done !" / "Wrong !" / "Time up ." in the step that
            ends it (xworld3d_task.py:456-482), "" afterwards;
          walls.json tasks: the command in the step whose teach() ran idle() (XWorldNav*.py), "Well done !" in the
            step that reached the goal (xworld_task.py:215-221), "" otherwise.
        One read-back of the small state fields for the whole batch; strings are built on the host."""
        if self.cfg.game != _abi.XW_GAME_XWORLD:
            raise RuntimeError("sentences(): xworld only")
        f = self.get_fields(SENTENCE_FIELDS)
        return [sentence_for_state(self._lib, self.cfg, self.catalog, self.cfg.env_id_offset + e,
                                   {k: f[k][e] for k in SENTENCE_FIELDS})
                for e in (range(self.n_envs) if envs is None else envs)]

    # ------------------------------------------------------------------ state access
    def get_field(self, name):
        out = self._field_buffer(name)
        self._check(self._lib.xw_get_field(self._h, name.encode(), out.ctypes.data, out.nbytes))
        return out

    def _field_buffer(self, name):
        n = self.n_envs
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
        elif name in ("agent_x", "agent_y", "facing", "task", "stage", "event", "action_success", "target_mask",
                      "aux0", "aux1", "aux2", "level"):
            out = np.zeros(n, np.uint8)
        elif name in ("win_len", "win_sum", "win_pos"):
            out = np.zeros((n, 5), np.uint8)
        elif name == "win_bits":
            out = np.zeros((n, 5, 7), np.uint32)
        elif name == "state":
            out = np.zeros((n, 4), np.float32)
        elif name in ("pos_x", "pos_y", "angle"):
            out = np.zeros(n, np.float32)
        else:
            out = np.zeros(n, np.int32)
        return out

    def get_fields(self, names):
        """Several fields with one device synchronisation (xw_get_fields)."""
        outs = {}
        for k in names:
            outs[k] = self._field_buffer(k)
        n = len(names)
        c_names = (C.c_char_p * n)(*[k.encode() for k in names])
        c_ptrs = (C.c_void_p * n)(*[outs[k].ctypes.data for k in names])
        c_bytes = (C.c_size_t * n)(*[outs[k].nbytes for k in names])
        self._check(self._lib.xw_get_fields(self._h, n, c_names, c_ptrs, c_bytes))
        return outs

    def set_field(self, name, value):
        cur = self.get_field(name)
        v = np.ascontiguousarray(value, cur.dtype).reshape(cur.shape)
        self._check(self._lib.xw_set_field(self._h, name.encode(), v.ctypes.data, v.nbytes))

    # ------------------------------------------------------------------ checkpoint / resume
    _XWORLD_STATE = ("grid", "agent_x", "agent_y", "facing", "task", "stage", "event", "action_success", "target_mask", "aux0",
                     "aux1", "aux2", "goal_x", "goal_y", "goal_icon", "goal_name", "steps_in_task", "num_steps", "episode",
                     "n_success", "n_failure", "success_steps", "minstd", "error")
    _CURRICULUM_STATE = ("level", "check_counter", "win_len", "win_sum", "win_pos", "win_bits")
    _FPV_STATE = ("goal_yaw", "goal_scale", "goal_offset")
    _RACE_STATE = ("pos_x", "pos_y", "angle", "steps", "state")

    def state_dict(self):
        """Everything a batch needs to continue bit for bit: the SoA state of every env (the episode counter and the
        minstd word are the RNG state: the Philox draws are counter-based), copied to the host.  The reference has no
        counterpart beyond --curriculum_stamp (xworld.cpp:92-135); SURVEY lists checkpoint/resume as an aux subsystem."""
        if self.cfg.game == _abi.XW_GAME_SIMPLE_GAME:
            raise RuntimeError("state_dict(): not available for simple_game")
        names = self._RACE_STATE if self.cfg.game == _abi.XW_GAME_SIMPLE_RACE else self._XWORLD_STATE + (
            self._CURRICULUM_STATE if self.cfg.curriculum != 0 else ()) + (
            self._FPV_STATE if self.cfg.visible_radius > 0 else ())
        sd = {k: self.get_field(k) for k in names}
        sd["_last_over"], sd["_last_reward"] = self._last_over.copy(), np.array(self._last_reward, np.float32)
        return sd

    def load_state_dict(self, sd):
        for k, v in sd.items():
            if not k.startswith("_"):
                self.set_field(k, v)
        self._last_over = np.array(sd["_last_over"], np.int32)
        self._last_reward = np.array(sd["_last_reward"], np.float32)

    def launch_count(self):
        return int(self._lib.xw_launch_count(self._h))

    def render_kernel(self):
        """0 generic, 1 plan compositor, 2 pipelined plan compositor, 3 sparse painter (xw_render_kernel)."""
        return int(self._lib.xw_render_kernel(self._h))
