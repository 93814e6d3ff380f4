/*
 * xworld_b200.h -- C ABI of the B200-native batched XWorld2D simulator.
 *
 * This is the drop-in boundary for ONE hot path of PaddlePaddle/XWorld:
 *   SimulatorInterface::take_actions  (simulator_interface.cpp:126-137)
 *     -> XWorldSimulator::take_action (games/xworld/xworld_simulator.cpp:200-265)
 *     -> XAgent::act / XMap::move_item (games/xworld/xworld/xitem.cpp:89-155, xmap.cpp:76-101)
 *     -> Teacher::teach + task reward rules (teacher.cpp:207-230, teaching_task.cpp:64-116,
 *        games/xworld3d/tasks/xworld3d_task.py:451-482, games/xworld/tasks/xworld_task.py:184-223)
 *     -> XWorldSimulator::get_screen    (xworld_simulator.cpp:278-307,508-545; xmap.cpp:125-146)
 * replayed for N environments at once by hand-written sm_100a CUDA kernels.
 *
 * Plain C: pointers and sizes only, no torch / STL types.  Every entry point names the
 * reference interface it replaces.  All functions return 0 on success, a negative
 * xw_status on failure (never abort -- the reference CHECK/LOG(FATAL)s instead,
 * e.g. xworld_simulator.cpp:254); xw_last_error() gives the message.
 *
 * Pointers called d_* are DEVICE pointers (cudaMalloc'ed / torch CUDA tensors),
 * h_* are HOST pointers.  `stream` is a cudaStream_t passed as void* (0 = default).
 */
#ifndef XWORLD_B200_H_
#define XWORLD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XW_ABI_VERSION 2
#define XW_MAX_GOALS 8   /* goals per map (XWorldNav uses 2..4, XWorldNav.py:31-32) */
#define XW_MAX_DIM 16    /* map side in cells (reference hard-codes 8, XWorldNav.py:10-11) */
#define XW_ICON_SIZE 64  /* XItem::item_size_, xitem.h:151 */

typedef enum {
    XW_OK = 0,
    XW_ERR_INVALID_ARG = -1,
    XW_ERR_CUDA = -2,
    XW_ERR_NO_DEVICE = -3,
    XW_ERR_INVALID_ACTION = -4, /* reference: CHECK_LT(action_idx, num_actions) aborts */
    XW_ERR_UNSUPPORTED = -5
} xw_status;

/* Which game the handle simulates (SimulatorInterface ctor, simulator_interface.cpp:37-85). */
typedef enum {
    XW_GAME_XWORLD = 0,      /* "xworld"      -> CUDA */
    XW_GAME_SIMPLE_GAME = 1, /* "simple_game" -> host (BASELINE config 1: CPU plumbing) */
    XW_GAME_SIMPLE_RACE = 2  /* "simple_race" -> CUDA, fp32 */
} xw_game;

/* --task_mode (xworld_simulator.cpp:34-37).  lang_acquisition: a teacher event ends the episode (game_over =
 * SUCCESS / DEAD / MAX_STEP, xworld_simulator.cpp:165-177).  one_channel: "each session has all tasks until the max
 * steps" (:192-193): game_over reports only --max_steps; the walls.json navigation tasks get their time-up rule
 * (xworld_task.py:205-210).  "interactive" (language-only sessions) is out of scope. */
typedef enum { XW_TASK_LANG_ACQUISITION = 0, XW_TASK_ONE_CHANNEL = 1 } xw_task_mode;

/* Teacher rule set = the task group of the conf json (teacher.cpp:70-99). */
typedef enum {
    XW_RULES_NAV3D = 0, /* confs/navigation2d.json: XWorld3DNav{Target,TargetNear,TargetBetween,
                           TargetDirection,TargetAvoid} (games/xworld3d/tasks/) */
    XW_RULES_NAV2D = 1  /* confs/walls.json: XWorldNav{Target,Near,ColorTarget,Between} +
                           XWorldRec* (games/xworld/tasks/) */
} xw_rules;

/* GameOverCode bitmask, simulator.h:42-48. */
enum { XW_ALIVE = 0, XW_MAX_STEP = 1, XW_DEAD = 2, XW_SUCCESS = 4, XW_LOST_LIFE = 8 };

/* Teacher event of the last step (TeachingEnvBuffer::event, simulator.h:267-291). */
enum { XW_EVENT_NONE = 0, XW_EVENT_CORRECT_GOAL = 1, XW_EVENT_WRONG_GOAL = 2, XW_EVENT_TIME_UP = 3 };

/* Task stage (Task::current_stage_, teaching_task.h:51-104). */
enum { XW_STAGE_IDLE = 0, XW_STAGE_NAVIGATION = 1, XW_STAGE_TERMINAL = 2 };

/* Task ids in conf-json order (TaskGroup::run_stage indexes task_list_, teaching_task.cpp:204-222). */
enum {
    XW_T3_TARGET = 0, XW_T3_NEAR = 1, XW_T3_BETWEEN = 2, XW_T3_DIRECTION = 3, XW_T3_AVOID = 4,
    XW_T2_TARGET = 0, XW_T2_NEAR = 1, XW_T2_COLOR_TARGET = 2, XW_T2_BETWEEN = 3
};

/* An action id that makes xw_step / xw_step_host / xw_step_hd leave that env alone: no move, no teacher call, no step
 * count, and its reward / game_over output slots are not written.  The reference has no such thing (one process per
 * env: an env that is not stepped is simply not called); a batch needs it to serve callers that step envs one at a
 * time (xworld_b200/wire.py).  The frame of such an env is re-rendered unchanged. */
#define XW_ACTION_NONE (-1)
/* Grid cell codes (one item per cell in navigation maps; XMap::item_ptr_cube_, xmap.h:95). */
enum { XW_CELL_EMPTY = 0, XW_CELL_BLOCK = 1, XW_CELL_AGENT = 2, XW_CELL_GOAL0 = 3 };

/* Icon catalog = what XWorldEnv.set_goal_subtrees builds from item_path
 * (games/xworld/maps/xworld_env.py:244-268) plus properties.txt colours (:88-94). */
typedef struct {
    int32_t n_icons;          /* icons in the atlas */
    int32_t brick_icon;       /* block/brick_1.jpg */
    int32_t agent_icon;       /* agent/robot_1.jpg */
    int32_t n_names;          /* goal class names in canonical (sorted) order */
    const int32_t* name_first;   /* [n_names+1] CSR offsets into name_icons */
    const int32_t* name_icons;   /* icon ids of each name's variants, sorted by path */
    const uint8_t* icon_colored; /* [n_icons] 1 iff properties.txt colour != "na" */
    const uint8_t* atlas64;   /* HOST, [n_icons][64][64][3] BGR bytes as cv::imread(path,1)
                                 returns them (xitem.cpp:38); parity is defined post-decode */
} xw_catalog;

/* Per-batch options = the reference's process-global gflags (simulator.cpp:21-27,
 * xworld_simulator.cpp:22-37, teacher.cpp:22-25, simulator_util.cpp:27) without the globals. */
typedef struct {
    int32_t abi_version;      /* XW_ABI_VERSION */
    int32_t game;             /* xw_game */
    /* ---- xworld ---- */
    int32_t height, width;    /* map size in cells (XWorldNav max_height/max_width) */
    int32_t n_goals, n_blocks;/* XWorldNav.py num_goals_seq / num_blocks_seq entries */
    int32_t rules;            /* xw_rules */
    int32_t out_h, out_w;     /* frame size; 0 = reference rule height*12 (xworld_simulator.cpp:52-61) */
    int32_t context;          /* --context (frames stacked, simulator.cpp:21) */
    int32_t max_steps;        /* --max_steps, 0 = off (simulator.h:68-74) */
    int32_t max_steps_factor; /* --max_steps_factor (xworld3d_task.py:38), default 10 */
    int32_t visible_radius;   /* --visible_radius (xworld_simulator.cpp:26): 0 = fully observed; an odd vr > 0 = the
                                 first-person view (xmap.cpp:148-200): a vr x vr cell window ahead of the agent, wall
                                 shadows, rotated to the agent's heading, frames of vr*(84/vr) pixels a side
                                 (xworld_simulator.cpp:62-68), six actions (xitem.cpp:82-86).  Clamped to the map side
                                 like the reference; an even value is an error (CHECK_EQ(vr % 2, 1), xmap.cpp:277) */
    int32_t auto_reset;       /* 0 = reference behaviour (caller resets) */
    int32_t simulator_seed;   /* --simulator_seed: env i seeds minstd_rand0 like the reference's
                                 i-th thread (simulator_util.cpp:38-55) */
    uint64_t seed;            /* Philox key for the draws the reference takes from Python's
                                 unseeded `random` (map generation, task idle stages) */
    int64_t env_id_offset;    /* global id of local env 0 (multi-GPU sharding) */
    /* ---- simple_game ---- */
    int32_t array_size;       /* --array_size (simple_game_simulator.cpp:18) */
    /* ---- simple_race (simple_race_simulator.cpp:17-26) ---- */
    int32_t track_type;       /* 0 straight, 1 circle */
    float track_width, track_length, track_radius;
    int32_t race_full_manouver;
    int32_t race_random;      /* --random (simple_race_simulator.cpp:24): start position / heading drawn per episode from the env's
                                 std::default_random_engine, seeded like the reference's simulator threads (simulator_seed) */
    int32_t difficulty;       /* 0 easy, 1 hard */
    float reward_scale;
    /* ---- xworld curriculum (XWorldNav.py:36-56, xworld_env.py:103-110) ---- */
    float curriculum;         /* --curriculum (teacher.cpp:25; read as a float, py_simulator.cpp:127): 0 = off = the
                                 last level's map every episode; > 0 = per-env level schedule: a (3+level)-sided
                                 world with num_goals_seq / num_blocks_seq[level] entities inside the 8x8 map, padded
                                 with bricks, one level up when the lowest per-task success rate over the last 200
                                 results reaches this threshold.  8x8 map only.  With the walls.json rules no task
                                 class ever records a result in lang_acquisition mode, so an env keeps its start level */
    int32_t curriculum_check_period; /* XWorldEnv.curriculum_check_period (xworld_env.py:58); 0 = the reference's 100 */
    int32_t start_level;      /* XWorldNav(start_level=...) (XWorldNav.py:8), 0..5 */
    /* ---- flags the reference reads per process (xworld_simulator.cpp:34-37, simulator.cpp:25) ---- */
    int32_t task_mode;        /* xw_task_mode: --task_mode.  The C++ flag default is "lang_acquisition" (= 0 here); the
                                 reference's PYTHON default is "one_channel" (py_simulator.cpp:128-130), which
                                 xworld_b200.Simulator.create mirrors */
    int32_t gray;             /* 1 = --color=false (the reference default, simulator.cpp:25): one-channel frames,
                                 cv::cvtColor(BGR2GRAY) after the resize (xworld_simulator.cpp:529-531) with the
                                 coefficients of the OpenCV 3.2.0 the reference pins (cmake/opencv.cmake:5-6):
                                 (B*1868 + G*9617 + R*4899 + 8192) >> 14.  0 = --color=true: planes B, G, R */
    int32_t reserved[3];
} xw_config;

typedef struct xw_sim xw_sim; /* opaque handle: one batch of n_envs environments on one GPU */

/* Fill *cfg with the reference's flag defaults. */
void xw_config_init(xw_config* cfg);

/* SimulatorInterface::SimulatorInterface(name) (simulator_interface.cpp:37-85) for a batch.
 * catalog may be NULL for simple_game / simple_race.  device < 0: current device. */
int xw_create(const xw_config* cfg, const xw_catalog* catalog, int32_t n_envs, int32_t device,
              xw_sim** out);
void xw_destroy(xw_sim* sim);
const char* xw_last_error(void);

/* SimulatorInterface::reset_game (simulator_interface.cpp:95-105): game reset, teacher reset +
 * first teach(), first frame.  d_mask: optional device u8[n_envs], reset only envs with mask!=0
 * (NULL = all). */
int xw_reset(xw_sim* sim, const uint8_t* d_mask, void* stream);

/* SimulatorInterface::take_actions (simulator_interface.cpp:126-137) for every env:
 * act_rep x take_action -> teach -> reward -> (render if d_frames != NULL).
 *   d_actions   i32[n_envs]   action ids (StatePacket "action", xworld_simulator.cpp:231)
 *   d_reward    f32[n_envs]   return value of take_actions
 *   d_game_over i32[n_envs]   SimulatorInterface::game_over() right after the step
 *   d_frames    u8[n_envs][context*C][out_h][out_w] or NULL (get_state()["screen"] bytes)
 * An action outside [0, xw_num_actions) leaves the env untouched (reward 0, game_over 0) and sets its error flag
 * (xw_error_flags); the reference CHECK-aborts the process (xworld_simulator.cpp:254).  This call is asynchronous and
 * cannot report it; the calls that hand results to the host (xw_step_host, xw_step_hd, xw_step_hd_async) return
 * XW_ERR_INVALID_ACTION when an env of that step was flagged -- the other envs were stepped normally. */
int xw_step(xw_sim* sim, const int32_t* d_actions, int32_t act_rep, float* d_reward,
            int32_t* d_game_over, uint8_t* d_frames, void* stream);

/* k_steps consecutive SimulatorInterface::take_actions calls of every env, no frames: d_actions, d_reward, d_game_over are
 * [k_steps][n_envs] (step-major); auto_reset applies between the steps.  Same results as k_steps xw_step calls with
 * d_frames == NULL.  simple_race runs them in ONE launch with the car in registers (open-loop action sequences:
 * evaluation roll-outs); xworld queues k_steps step (+ reset) launches.  No reference counterpart: the reference's
 * simulation_loop (simulator_interface.cpp:361-435) is one request per step by construction. */
int xw_step_seq(xw_sim* sim, const int32_t* d_actions, int32_t k_steps, int32_t act_rep, float* d_reward,
                int32_t* d_game_over, void* stream);

/* GameSimulator::make_context_screens -> XWorldSimulator::get_screen (simulator.cpp:62-85,
 * xworld_simulator.cpp:278-285): render the current state of every env. */
int xw_render(xw_sim* sim, uint8_t* d_frames, void* stream);

/* Same call with HOST buffers (pinned or pageable); copies in/out on the handle's stream and
 * synchronises.  h_frames may be NULL.  This is the end-to-end path bench.py times. */
int xw_step_host(xw_sim* sim, const int32_t* h_actions, int32_t act_rep, float* h_reward,
                 int32_t* h_game_over, uint8_t* h_frames);
int xw_reset_host(xw_sim* sim, const uint8_t* h_mask, uint8_t* h_frames);
/* Host actions / reward / game_over, frames rendered into a DEVICE buffer (NULL = no render): the
 * call a trainer with a co-located GPU learner makes.  Copies + synchronises like xw_step_host. */
int xw_step_hd(xw_sim* sim, const int32_t* h_actions, int32_t act_rep, float* h_reward,
               int32_t* h_game_over, uint8_t* d_frames);

/* xw_step_hd without the wait for the frames: returns as soon as reward / game_over are on the host, while the render
 * kernel of the step may still be running on the handle's stream (the next call's step kernel queues behind it, so the
 * GPU never waits for the host between steps).  Anything that reads d_frames must be ordered after the render:
 * xw_wait_frames makes `stream` wait for it (event), xw_sync the host.  A caller that lets a consumer on another stream
 * read the frames of step t while step t+1 is issued should alternate between two d_frames buffers.  No reference
 * counterpart (the reference renders on the CPU inside take_actions). */
int xw_step_hd_async(xw_sim* sim, const int32_t* h_actions, int32_t act_rep, float* h_reward,
                     int32_t* h_game_over, uint8_t* d_frames);
int xw_wait_frames(xw_sim* sim, void* stream);
int xw_sync(xw_sim* sim);
/* get_num_actions / get_screen_out_dimensions / get_num_steps / get_lives
 * (simulator_interface.h:52-63). */
int32_t xw_num_envs(const xw_sim* sim);
int32_t xw_num_actions(const xw_sim* sim);
int xw_screen_dims(const xw_sim* sim, int32_t* h, int32_t* w, int32_t* c, int32_t* context);
size_t xw_frame_bytes(const xw_sim* sim); /* context*C*out_h*out_w */
int xw_num_steps(xw_sim* sim, int64_t* h_num_steps /* [n_envs] host */);

/* State fields by name, copied to host (tests, checkpointing, get_extra_info).  Fields:
 * "grid" u8[n][H*W], "agent_x","agent_y","facing","task","stage","event","action_success",
 * "target_mask","aux0","aux1","aux2" u8[n]; "goal_x","goal_y" u8[n][XW_MAX_GOALS];
 * "goal_icon" i32[n][XW_MAX_GOALS]; "steps_in_task","num_steps","episode","n_success",
 * "n_failure","success_steps","minstd","error" i32[n].
 * First-person view (cfg.visible_radius > 0): "goal_yaw" u16[n][XW_MAX_GOALS] (yaw = 4 * 1.5707963 * idx / 4096),
 * "goal_scale","goal_offset" f64[n][XW_MAX_GOALS] (Entity.yaw / scale / offset, xworld_env.py:211-223); setting any "goal_*"
 * field re-warps the cached goal icons.  "facing" is the agent's heading: 0 right, 1 down, 2 left, 3 up.
 * Curriculum (only when cfg.curriculum > 0): "level" u8[n] (XWorldEnv.dump_curriculum_progress, xworld_env.py:62-63),
 * "check_counter" i32[n], "win_len","win_sum" u8[n][5] (length / successes of each task class's result window),
 * "win_pos" u8[n][5] and "win_bits" u32[n][5][7] (the windows themselves: with these a checkpoint restores everything).
 * Race: "pos_x","pos_y","angle" f32[n], "steps" i32[n], "state" f32[n][4]. */
int xw_get_field(xw_sim* sim, const char* name, void* h_out, size_t bytes);
/* The same for several fields with one device synchronisation (e.g. the ten fields a sentence is composed from). */
int xw_get_fields(xw_sim* sim, int32_t n_fields, const char* const* names, void* const* h_out, const size_t* bytes);

/* SimulatorInterface::get_world_dimensions (simulator_interface.cpp:163-167 -> xworld_simulator.cpp:100-104): X = width,
 * Y = height, Z = 0 for xworld; the values are left untouched for the other games, as in the reference. */
int xw_world_dimensions(const xw_sim* sim, double* X, double* Y, double* Z);

/* XWorldSimulator::get_extra_info (xworld_simulator.cpp:495-504) for one env:
 *   "<id>|task:<teacher sentence type>,event:<event>,height:<h>,width:<w>"
 * the string py_simulator's get_state() splits into its "task" / "event" / "height" / "width" keys (py_simulator.cpp:276-283).
 * <id> is the env's global id where the reference prints its process id; <h>, <w> are the world's own side (the level's,
 * under --curriculum).  Returns the length, or a negative status.  Empty for the other games, as in the reference. */
int xw_extra_info(xw_sim* sim, int32_t env, char* buf, size_t cap);

/* SimulatorInterface::teacher_report_task_performance (simulator_interface.cpp:149-153 -> Teacher::report_task_performance,
 * teacher.cpp:175-200): successes, failures and the steps spent in successful episodes per task class, summed over the batch
 * since the handle was created.  Arrays of `cap` >= 5 entries; names (may be NULL) receives static strings.  Returns the number
 * of task classes.  (Formatting the "=== S(S)/F(F) -> rate@steps" lines is the caller's: include/xworld_b200.hpp does it.) */
int xw_task_performance(xw_sim* sim, int64_t* successes, int64_t* failures, int64_t* success_steps, int32_t cap,
                        const char** names);

/* Per-env error flags: 0, or XW_ERR_INVALID_ACTION since the env's last reset.  Copies them to h_flags[n_envs] (may be
 * NULL) and returns how many envs are flagged (< 0: a CUDA error).  Synchronises the device. */
int32_t xw_error_flags(xw_sim* sim, int32_t* h_flags);
int xw_set_field(xw_sim* sim, const char* name, const void* h_in, size_t bytes);

/* ---- teacher language channel, host side ------------------------------------------------------
 * Replaces CFG.generate (python/context_free_grammar.py:166-188) over the grammars of the nine navigation tasks
 * (games/xworld3d/tasks/XWorld3DNavTarget*.py, games/xworld/tasks/XWorldNav*.py: _define_grammar + the _bind calls
 * of idle()).  The reference draws productions from Python's unseeded `random`; here draw i of a sentence is
 * Philox(seed, env_id, episode, salt, i), so a sentence is a pure function of the query.  No device access: the
 * caller reads the slots from the env's state ("task", "aux0", "aux1", "goal_name", "episode"; Simulator.sentences()
 * in xworld_b200/simulator.py does it for a whole batch).  Returns the sentence length (0 = no sentence for this
 * query) or a negative status. */
enum { XW_SENT_START = 0, XW_SENT_CORRECT = 1, XW_SENT_WRONG = 2, XW_SENT_TIMEUP = 3 };
typedef struct xw_sentence_query {
    int32_t rules;      /* XW_RULES_NAV3D | XW_RULES_NAV2D */
    int32_t task;       /* XW_T3_* | XW_T2_* */
    int32_t kind;       /* XW_SENT_*: `S -> start | correct/finish | wrong | timeup` */
    int32_t direction;  /* XW_T3_DIRECTION: 1 front, 2 behind, 3 left, 4 right (the engine's "aux1") */
    const char* name1;  /* G / G1 / O: goal class name */
    const char* name2;  /* G2 / T (Between tasks) */
    const char* color;  /* C (XW_T2_COLOR_TARGET) */
    uint64_t seed;      /* xw_config.seed */
    int64_t env_id;     /* global env id (env_id_offset + index) */
    uint32_t episode;   /* the env's "episode" field */
    uint32_t salt;      /* 0, or the env's step count when an episode issues several commands (walls.json tasks) */
} xw_sentence_query;
int xw_sentence_compose(const xw_sentence_query* q, char* buf, size_t cap);

/* ---- SimulatorServer / SimulatorClient wire format, host side, no sockets ------------------------------
 * The reference runs one simulator process per env; the trainer's SimulatorServer (simulator_interface.cpp:165-320)
 * sends each a length-prefixed util::BinaryBuffer message (simulator_communication.h:34-76,222-240; memory_util.h:83-115)
 * whose packets are StatePacket::encode (data_packet.h:315-333, data_packet.cpp:137-174), and SimulatorClient::
 * simulation_loop (simulator_interface.cpp:361-435) answers.  These functions parse those requests and build those
 * replies byte for byte, so that N such connections can be served from the N envs of one batch (xworld_b200/wire.py).
 * Encoders return the number of bytes the message takes and write it only if it fits in `cap` (out may be NULL to
 * size a buffer); replies and requests are FRAMED: 8-byte body size, then the body. */
#define XW_WIRE_MAX_FIELDS 8
typedef struct xw_wire_field {     /* one StatePacket entry = key + StateBuffer; a NULL pointer = that part is absent */
    const char* key;
    const float* reals;   uint64_t n_reals;   /* std::vector<float>   (flag bit 1) */
    const uint8_t* pixels; uint64_t n_pixels; /* std::vector<uint8_t> (flag bit 2) */
    const int32_t* ids;   uint64_t n_ids;     /* std::vector<int>     (flag bit 4) */
    const char* str;                          /* std::string          (flag bit 8) */
} xw_wire_field;
typedef struct xw_wire_request {
    const char* cmd;        /* "reset" | "take_actions" | "get_state" | "report_perf" | "get_extra_info" | "stop" */
    int32_t act_rep;        /* take_actions */
    int32_t show_screen;    /* take_actions */
    float reward;           /* get_state */
    int32_t n_fields;       /* take_actions: the action packet ("action" ids, "pred_sentence" str) */
    xw_wire_field fields[XW_WIRE_MAX_FIELDS];
} xw_wire_request;
/* StatePacket::encode / decode.  Decoded pointers point INTO `in` (keys and strings are NUL-terminated on the wire);
 * float / int arrays may be unaligned there: memcpy them out. */
int64_t xw_wire_encode_packet(const xw_wire_field* fields, int32_t n_fields, uint8_t* out, size_t cap);
int xw_wire_decode_packet(const uint8_t* in, size_t len, xw_wire_field* fields, int32_t max_fields, int32_t* n_fields,
                          size_t* consumed);
/* A request body (after the 8-byte size), as CommServer::call_remote_func composes it. */
int xw_wire_parse_request(const uint8_t* body, size_t len, xw_wire_request* req);
/* The server side of the same, for tests and for trainers written against this library. */
int64_t xw_wire_compose_request(const char* cmd, const xw_wire_field* fields, int32_t n_fields, int32_t act_rep, int32_t show_screen,
                        float reward, uint8_t* out, size_t cap);
/* SimulatorClient::reset_game / take_actions / get_state / get_extra_info replies (simulator_interface.cpp:385-435);
 * xw_wire_reply_text(cmd, NULL) is the echo that answers "report_perf". */
int64_t xw_wire_reply_reset(int32_t num_actions, int32_t game_over, int32_t lives, uint64_t height, uint64_t width,
                            uint64_t channels, double X, double Y, double Z, uint8_t* out, size_t cap);
int64_t xw_wire_reply_take_actions(float reward, int64_t num_steps, int32_t game_over, int32_t lives, int32_t action_success,
                                   const char* last_action, uint8_t* out, size_t cap);
int64_t xw_wire_reply_get_state(const xw_wire_field* fields, int32_t n_fields, uint8_t* out, size_t cap);
int64_t xw_wire_reply_text(const char* cmd, const char* text, uint8_t* out, size_t cap);
/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t xw_launch_count(const xw_sim* sim);
/* Which render kernel the handle uses (diagnostics, tests): 0 = generic per-byte kernel, 1 = plan compositor with
 * one frame buffer per warp group, 2 = pipelined plan compositor, 3 = sparse painter (the default when the frame
 * geometry allows; XW_RENDER_MODE=sb|pipe selects the others); first-person view: 4 = per-pixel kernel (any frame size),
 * 5 = shared-memory frame kernel (frame width % 4 == 0 and frame bytes % 16 == 0), 6 = the same by whole cell blocks
 * (geometries where every frame pixel takes its taps from one cell, e.g. visible_radius 7 on an 11x11 map);
 * -1 = the game has no renderer.  No reference
 * counterpart: the reference has one OpenCV code path (xworld_simulator.cpp:278-307). */
int32_t xw_render_kernel(const xw_sim* sim);
/* CUDA-event timing of the render kernel alone: average ms over the launches since the last
 * call with reset != 0.  Returns <0 if timing is disabled.  xw_enable_timing(sim, 1) first. */
int xw_enable_timing(xw_sim* sim, int32_t on);
double xw_render_ms(xw_sim* sim, int32_t reset);
/* The same for the launches in front of the render kernel: k_step + the auto-reset launch (+ the first-person goal warp). */
double xw_step_reset_ms(xw_sim* sim, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* XWORLD_B200_H_ */
