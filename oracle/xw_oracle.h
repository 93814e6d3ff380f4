/*
 * xw_oracle.h -- CPU oracle for the XWorld2D hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under xworld_b200/ links or calls it.
 *
 * Every function restates one piece of /root/reference (file:line in the .c file).
 */
#ifndef XW_ORACLE_H_
#define XW_ORACLE_H_
#include <stdint.h>
#include "../include/xworld_b200.h" /* xw_config / xw_catalog / enums only (the interface) */

#ifdef __cplusplus
extern "C" {
#endif

/* Philox substream ids: one per reference call site that draws from Python's `random`. */
enum {
    XO_SITE_NAMES = 1,     /* XWorldNav._configure: random.shuffle(goal_names)            XWorldNav.py:60 */
    XO_SITE_MAZE = 2,      /* maze2d.dfs: random.shuffle(moves), 3 draws per visited node  maze2d.py:95 */
    XO_SITE_BLOCKS = 3,    /* __instantiate_entities: random.shuffle(blocks)               xworld_env.py:424 */
    XO_SITE_GOAL_LOC = 4,  /* set_property: loc = choice(available_grids), index = goal#   xworld_env.py:190 */
    XO_SITE_GOAL_ASSET = 5,/* set_property: asset_path = choice(items[type][name])         xworld_env.py:198 */
    XO_SITE_AGENT_LOC = 6, /* set_property on the agent                                    xworld_env.py:190 */
    XO_SITE_TASK_A = 7,    /* task idle: first choice (sel_goal / tile)                    XWorld3DNav*.py idle */
    XO_SITE_TASK_B = 8,    /* task idle: second choice (referent / empty grid e)           */
    XO_SITE_TASK_SHUF = 9, /* task idle: random.shuffle(goals); g1, g2 = goals[:2]         */
    XO_SITE_TASK_AGENT = 10,/* task idle: agent.loc = choice(new_a)                         */
    /* 11 = the sentence channel (xworld_b200/csrc/xw_sentence.hpp) */
    XO_SITE_AGENT_YAW = 12,/* set_property, --visible_radius > 0: agent yaw = choice(range(-1,3)) * PI_2   xworld_env.py:208-210 */
    XO_SITE_GOAL_POSE = 13 /* set_property, --visible_radius > 0: goal yaw / scale / offset = uniform(...)  xworld_env.py:211-223;
                              index = 4 * goal# + {0 yaw, 1 scale, 2 offset} */
};
#define XO_YAW_STEPS 4096 /* goal yaws are drawn on a 4096-point grid of [0, 4 * PI_2): see xo_goal_pose */

typedef struct {
    int32_t H, W;
    uint8_t grid[XW_MAX_DIM * XW_MAX_DIM];
    int32_t agent_x, agent_y;
    double agent_yaw;
    int32_t n_goals;
    int32_t goal_x[XW_MAX_GOALS], goal_y[XW_MAX_GOALS], goal_icon[XW_MAX_GOALS], goal_name[XW_MAX_GOALS];
    int32_t task, stage, event, action_success;
    int32_t target_mask, aux0, aux1, aux2;
    int32_t steps_in_task;
    int64_t num_steps;
    int32_t episode;
    int32_t n_success, n_failure, success_steps;
    uint32_t minstd;
    int32_t error;
    int64_t env_gid;
    /* curriculum (cfg->curriculum > 0): XWorldEnv.current_level / curriculum_check_counter, the world's side
     * (XWorldEnv.height == width) and each task class's success_seq (xworld3d_task.py:67,129-133) */
    int32_t level, dim, check_counter;
    int32_t seq_len[5];
    uint8_t seq[5][200];
    /* --visible_radius > 0: Entity.yaw / scale / offset of each goal (xworld_env.py:211-223); agent_yaw above is the
     * agent's.  Defaults (fully observed): yaw 1.5707963, scale 1, offset 0 (xworld_env.py:41-42). */
    double goal_yaw[XW_MAX_GOALS], goal_scale[XW_MAX_GOALS], goal_offset[XW_MAX_GOALS];
} xo_env;

/* ---- RNG ---- */
void xo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint32_t xo_draw(uint64_t seed, int64_t env_gid, uint32_t episode, uint32_t attempt, uint32_t site,
                 uint32_t index);
uint32_t xo_randbelow(uint32_t u, uint32_t n);
uint64_t xo_std_hash_bytes(const void* p, uint64_t len); /* libstdc++ std::hash<std::string> */
uint32_t xo_minstd_seed_for_thread(int32_t simulator_seed, int32_t thread_no);
int32_t xo_get_rand_ind(uint32_t* minstd, int32_t size); /* simulator_util.cpp:66-73 */

/* ---- map pieces (exposed for unit tests) ---- */
void xo_maze(uint64_t seed, int64_t env_gid, uint32_t episode, uint32_t attempt, int D, char* maze /* D*D */);

/* ---- env ---- */
void xo_env_init(const xw_config* cfg, xo_env* e, int64_t env_gid);
int xo_reset(const xw_config* cfg, const xw_catalog* cat, xo_env* e);
int xo_step(const xw_config* cfg, const xw_catalog* cat, xo_env* e, int32_t action, int32_t act_rep,
            float* reward, int32_t* game_over);
/* teacher stage run once on an explicit state (used to pin rules against the reference python) */
int xo_teach(const xw_config* cfg, const xw_catalog* cat, xo_env* e, int32_t collided_cell_code,
             double* reward);

/* ---- first-person view (xw_oracle_fpv.c) ---- */
/* cv::getRotationMatrix2D (OpenCV imgproc/imgwarp.cpp), center given as Point2f */
void xo_rotation_matrix(float cx, float cy, double angle_deg, double scale, double M[6]);
/* cv::warpAffine, INTER_LINEAR, BORDER_CONSTANT, 8UC3, forward matrix M (inverted inside, as OpenCV does) */
void xo_warp_affine_8uc3(const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw, const double M[6],
                         const uint8_t border[3]);
/* XItem::get_item_image's transform of one 64x64 icon (xitem.cpp:47-60) */
void xo_item_image(const uint8_t* icon64, double yaw, double scale, double offset, uint8_t* out64);
/* XItem::get_item_facing_dir (xitem.cpp:65-78): 0 right, 1 down, 2 left, 3 up */
int xo_facing_dir(double yaw);
/* XMap::image_masking (xmap.cpp:273-362): ROI (in cells of the padded map) and the vr*vr shadow flags */
void xo_image_masking(const xo_env* e, int vr, int* x_st, int* y_st, uint8_t* shadow);
/* min(--visible_radius, max(h, w)) (xworld_simulator.cpp:63-64) */
int xo_visible_radius(const xw_config* cfg);
/* yaw / scale / offset of goal k as set_property draws them (xworld_env.py:211-223) */
void xo_goal_pose(uint64_t seed, int64_t env_gid, uint32_t episode, uint32_t attempt, int k, double* yaw, double* scale,
                  double* offset);
void xo_render_fpv(const xw_config* cfg, const xw_catalog* cat, const xo_env* e, uint8_t* frame_out);

/* ---- render ---- */
void xo_resize_tables(int src, int dst, int32_t* ofs, int16_t* a0, int16_t* a1);
void xo_resize_linear_8uc3(const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw);
/* The reference pipeline: white canvas -> per-item icon blit -> HWC->CHW -> CHW->HWC -> resize ->
 * HWC->CHW (B,G,R planes).  scratch must hold 3 * (H*64)*(W*64)*3 bytes (or NULL: malloc). */
void xo_render(const xw_config* cfg, const xw_catalog* cat, const xo_env* e, uint8_t* frame_out);
void xo_frame_dims(const xw_config* cfg, int* oh, int* ow);

/* ---- batch helpers (OpenMP over envs) for the CPU baseline ---- */
int xo_batch_reset(const xw_config* cfg, const xw_catalog* cat, xo_env* envs, int n, int threads);
int xo_batch_step(const xw_config* cfg, const xw_catalog* cat, xo_env* envs, int n, const int32_t* actions,
                  int act_rep, float* reward, int32_t* game_over, uint8_t* frames /* or NULL */, int threads);
int xo_sizeof_env(void);

/* ---- simple_game (games/simple_game/simple_game_simulator.cpp) ---- */
typedef struct { int32_t array_size, cur_pos; float rewards[64]; uint8_t state[64]; } xo_simple_game;
void xo_sg_reset(xo_simple_game* g, int array_size);
float xo_sg_act(xo_simple_game* g, int action);
int xo_sg_game_over(const xo_simple_game* g);

/* ---- simple_race (games/simple_race/simple_race_simulator.cpp) ---- */
typedef struct { float pos_x, pos_y, angle; int32_t steps; uint32_t minstd; /* --random: the env's (thread's) engine */ } xo_race;
float xo_rand_range_val(uint32_t* minstd, float upper); /* util::get_rand_range_val, simulator_util.cpp:57-64 */
void xo_race_reset(const xw_config* cfg, xo_race* r);
float xo_race_act(const xw_config* cfg, xo_race* r, int action_index, float state[4], int32_t* game_over);
float xo_race_take_actions(const xw_config* cfg, xo_race* r, int action_index, int act_rep, float state[4], int32_t* game_over);
/* n envs x steps on one thread (bench.py's CPU leg): actions[steps][n]; a finished game is reset; returns the reward sum */
double xo_race_batch(const xw_config* cfg, xo_race* envs, int n, const int32_t* actions, int steps);

#ifdef __cplusplus
}
#endif
#endif
