"""ctypes mirror of include/xworld_b200.h (the C ABI) and the loader of the in-tree CUDA library.

The product path has no CPU fallback: if libxworld_b200.so is missing, `load()` raises.
"""
import ctypes as C
import os

XW_ABI_VERSION = 2
XW_MAX_GOALS = 8
XW_ERR_INVALID_ARG, XW_ERR_CUDA, XW_ERR_NO_DEVICE, XW_ERR_INVALID_ACTION, XW_ERR_UNSUPPORTED = -1, -2, -3, -4, -5
XW_ACTION_NONE = -1
XW_WIRE_MAX_FIELDS = 8
XW_MAX_DIM = 16
XW_ICON_SIZE = 64

XW_GAME_XWORLD, XW_GAME_SIMPLE_GAME, XW_GAME_SIMPLE_RACE = 0, 1, 2
XW_RULES_NAV3D, XW_RULES_NAV2D = 0, 1
XW_TASK_LANG_ACQUISITION, XW_TASK_ONE_CHANNEL = 0, 1
XW_ALIVE, XW_MAX_STEP, XW_DEAD, XW_SUCCESS, XW_LOST_LIFE = 0, 1, 2, 4, 8
XW_EVENT_NONE, XW_EVENT_CORRECT_GOAL, XW_EVENT_WRONG_GOAL, XW_EVENT_TIME_UP = 0, 1, 2, 3
XW_STAGE_IDLE, XW_STAGE_NAVIGATION, XW_STAGE_TERMINAL = 0, 1, 2
XW_CELL_EMPTY, XW_CELL_BLOCK, XW_CELL_AGENT, XW_CELL_GOAL0 = 0, 1, 2, 3


class XwCatalog(C.Structure):
    _fields_ = [
        ("n_icons", C.c_int32),
        ("brick_icon", C.c_int32),
        ("agent_icon", C.c_int32),
        ("n_names", C.c_int32),
        ("name_first", C.POINTER(C.c_int32)),
        ("name_icons", C.POINTER(C.c_int32)),
        ("icon_colored", C.POINTER(C.c_uint8)),
        ("atlas64", C.POINTER(C.c_uint8)),
    ]


class XwConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("game", C.c_int32),
        ("height", C.c_int32),
        ("width", C.c_int32),
        ("n_goals", C.c_int32),
        ("n_blocks", C.c_int32),
        ("rules", C.c_int32),
        ("out_h", C.c_int32),
        ("out_w", C.c_int32),
        ("context", C.c_int32),
        ("max_steps", C.c_int32),
        ("max_steps_factor", C.c_int32),
        ("visible_radius", C.c_int32),
        ("auto_reset", C.c_int32),
        ("simulator_seed", C.c_int32),
        ("seed", C.c_uint64),
        ("env_id_offset", C.c_int64),
        ("array_size", C.c_int32),
        ("track_type", C.c_int32),
        ("track_width", C.c_float),
        ("track_length", C.c_float),
        ("track_radius", C.c_float),
        ("race_full_manouver", C.c_int32),
        ("race_random", C.c_int32),
        ("difficulty", C.c_int32),
        ("reward_scale", C.c_float),
        ("curriculum", C.c_float),
        ("curriculum_check_period", C.c_int32),
        ("start_level", C.c_int32),
        ("task_mode", C.c_int32),
        ("gray", C.c_int32),
        ("reserved", C.c_int32 * 3),
    ]


def default_config(**kw):
    """Reference flag defaults (simulator.cpp:21-27, xworld_simulator.cpp:22-37,
    simple_game_simulator.cpp:18, simple_race_simulator.cpp:17-26)."""
    cfg = XwConfig()
    cfg.abi_version = XW_ABI_VERSION
    cfg.game = XW_GAME_XWORLD
    cfg.height = cfg.width = 8
    cfg.n_goals, cfg.n_blocks = 4, 16
    cfg.rules = XW_RULES_NAV3D
    cfg.context = 1
    cfg.max_steps = 0
    cfg.max_steps_factor = 10
    cfg.array_size = 6
    cfg.track_width, cfg.track_length, cfg.track_radius = 20.0, 100.0, 30.0
    cfg.reward_scale = 1.0
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise TypeError("unknown xw_config field: %s" % k)
        setattr(cfg, k, v)
    return cfg


_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libxworld_b200.so")

XW_SENT_START, XW_SENT_CORRECT, XW_SENT_WRONG, XW_SENT_TIMEUP = 0, 1, 2, 3


class XwSentenceQuery(C.Structure):
    _fields_ = [("rules", C.c_int32), ("task", C.c_int32), ("kind", C.c_int32), ("direction", C.c_int32),
                ("name1", C.c_char_p), ("name2", C.c_char_p), ("color", C.c_char_p),
                ("seed", C.c_uint64), ("env_id", C.c_int64), ("episode", C.c_uint32), ("salt", C.c_uint32)]


class XwWireField(C.Structure):
    _fields_ = [("key", C.c_char_p), ("reals", C.POINTER(C.c_float)), ("n_reals", C.c_uint64),
                ("pixels", C.POINTER(C.c_uint8)), ("n_pixels", C.c_uint64), ("ids", C.POINTER(C.c_int32)), ("n_ids", C.c_uint64),
                ("str", C.c_char_p)]


class XwWireRequest(C.Structure):
    _fields_ = [("cmd", C.c_char_p), ("act_rep", C.c_int32), ("show_screen", C.c_int32), ("reward", C.c_float),
                ("n_fields", C.c_int32), ("fields", XwWireField * XW_WIRE_MAX_FIELDS)]


# every symbol include/xworld_b200.h declares
SYMBOLS = [
    "xw_config_init", "xw_create", "xw_destroy", "xw_last_error", "xw_reset", "xw_step", "xw_step_seq", "xw_render",
    "xw_step_host", "xw_reset_host", "xw_step_hd", "xw_step_hd_async", "xw_wait_frames", "xw_sync", "xw_num_envs", "xw_num_actions", "xw_screen_dims",
    "xw_frame_bytes", "xw_num_steps", "xw_get_field", "xw_set_field", "xw_error_flags", "xw_get_fields",
    "xw_world_dimensions", "xw_extra_info", "xw_task_performance", "xw_launch_count", "xw_render_kernel", "xw_sentence_compose",
    "xw_enable_timing", "xw_render_ms", "xw_step_reset_ms",
    "xw_wire_encode_packet", "xw_wire_decode_packet", "xw_wire_parse_request", "xw_wire_compose_request", "xw_wire_reply_reset",
    "xw_wire_reply_take_actions", "xw_wire_reply_get_state", "xw_wire_reply_text",
]


def load():
    """dlopen the in-tree CUDA library and declare the prototypes.  Fails loudly if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "xworld_b200: %s not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
            "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.xw_config_init.argtypes = [C.POINTER(XwConfig)]
    lib.xw_config_init.restype = None
    lib.xw_create.argtypes = [C.POINTER(XwConfig), C.POINTER(XwCatalog), i32, i32, C.POINTER(vp)]
    lib.xw_create.restype = C.c_int
    lib.xw_destroy.argtypes = [vp]
    lib.xw_destroy.restype = None
    lib.xw_last_error.argtypes = []
    lib.xw_last_error.restype = C.c_char_p
    lib.xw_reset.argtypes = [vp, vp, vp]
    lib.xw_reset.restype = C.c_int
    lib.xw_step.argtypes = [vp, vp, i32, vp, vp, vp, vp]
    lib.xw_step.restype = C.c_int
    lib.xw_step_seq.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    lib.xw_step_seq.restype = C.c_int
    lib.xw_render.argtypes = [vp, vp, vp]
    lib.xw_render.restype = C.c_int
    lib.xw_step_host.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.xw_step_host.restype = C.c_int
    lib.xw_step_hd.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.xw_step_hd.restype = C.c_int
    lib.xw_step_hd_async.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.xw_step_hd_async.restype = C.c_int
    lib.xw_wait_frames.argtypes = [vp, vp]
    lib.xw_wait_frames.restype = C.c_int
    lib.xw_sync.argtypes = [vp]
    lib.xw_sync.restype = C.c_int
    lib.xw_reset_host.argtypes = [vp, vp, vp]
    lib.xw_reset_host.restype = C.c_int
    lib.xw_num_envs.argtypes = [vp]
    lib.xw_num_envs.restype = i32
    lib.xw_num_actions.argtypes = [vp]
    lib.xw_num_actions.restype = i32
    lib.xw_screen_dims.argtypes = [vp] + [C.POINTER(i32)] * 4
    lib.xw_screen_dims.restype = C.c_int
    lib.xw_frame_bytes.argtypes = [vp]
    lib.xw_frame_bytes.restype = C.c_size_t
    lib.xw_num_steps.argtypes = [vp, vp]
    lib.xw_num_steps.restype = C.c_int
    lib.xw_get_field.argtypes = [vp, C.c_char_p, vp, C.c_size_t]
    lib.xw_get_field.restype = C.c_int
    lib.xw_set_field.argtypes = [vp, C.c_char_p, vp, C.c_size_t]
    lib.xw_set_field.restype = C.c_int
    lib.xw_get_fields.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.xw_get_fields.restype = C.c_int
    lib.xw_world_dimensions.argtypes = [vp] + [C.POINTER(C.c_double)] * 3
    lib.xw_world_dimensions.restype = C.c_int
    lib.xw_extra_info.argtypes = [vp, i32, C.c_char_p, C.c_size_t]
    lib.xw_extra_info.restype = C.c_int
    lib.xw_task_performance.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), i32, C.POINTER(C.c_char_p)]
    lib.xw_task_performance.restype = C.c_int
    lib.xw_error_flags.argtypes = [vp, vp]
    lib.xw_error_flags.restype = i32
    lib.xw_launch_count.argtypes = [vp]
    lib.xw_sentence_compose.argtypes = [C.POINTER(XwSentenceQuery), C.c_char_p, C.c_size_t]
    lib.xw_sentence_compose.restype = C.c_int
    lib.xw_render_kernel.argtypes = [vp]
    lib.xw_render_kernel.restype = C.c_int32
    lib.xw_launch_count.restype = i64
    lib.xw_enable_timing.argtypes = [vp, i32]
    lib.xw_enable_timing.restype = C.c_int
    lib.xw_render_ms.argtypes = [vp, i32]
    lib.xw_render_ms.restype = C.c_double
    lib.xw_step_reset_ms.argtypes = [vp, i32]
    lib.xw_step_reset_ms.restype = C.c_double
    u8p, sz, wf = C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(XwWireField)
    lib.xw_wire_encode_packet.argtypes = [wf, i32, u8p, sz]
    lib.xw_wire_encode_packet.restype = i64
    lib.xw_wire_decode_packet.argtypes = [u8p, sz, wf, i32, C.POINTER(i32), C.POINTER(sz)]
    lib.xw_wire_decode_packet.restype = C.c_int
    lib.xw_wire_parse_request.argtypes = [u8p, sz, C.POINTER(XwWireRequest)]
    lib.xw_wire_parse_request.restype = C.c_int
    lib.xw_wire_compose_request.argtypes = [C.c_char_p, wf, i32, i32, i32, C.c_float, u8p, sz]
    lib.xw_wire_compose_request.restype = i64
    lib.xw_wire_reply_reset.argtypes = [i32, i32, i32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_double, u8p, sz]
    lib.xw_wire_reply_reset.restype = i64
    lib.xw_wire_reply_take_actions.argtypes = [C.c_float, i64, i32, i32, i32, C.c_char_p, u8p, sz]
    lib.xw_wire_reply_take_actions.restype = i64
    lib.xw_wire_reply_get_state.argtypes = [wf, i32, u8p, sz]
    lib.xw_wire_reply_get_state.restype = i64
    lib.xw_wire_reply_text.argtypes = [C.c_char_p, C.c_char_p, u8p, sz]
    lib.xw_wire_reply_text.restype = i64
    _LIB = lib
    return lib
