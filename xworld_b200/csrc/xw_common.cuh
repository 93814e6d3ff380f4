// xw_common.cuh -- shared definitions of the batched XWorld2D engine (device state, RNG).
//
// State lives in HBM as a structure of arrays indexed by env (one warp lane steps one env, so
// every scalar field is a coalesced load); the item grid is an [env][cell] byte matrix with a
// 16-byte-multiple row stride because both its consumers want a whole env's cells at once
// (render) or one data-dependent cell (step), neither of which a [cell][env] layout coalesces.
#pragma once
#include <stdint.h>

#include "../../include/xworld_b200.h"

#if defined(__CUDACC__)
#define XW_HD __host__ __device__ __forceinline__
#else
#define XW_HD inline
#endif

// Philox substreams: one per reference call site that draws from Python's unseeded `random`
// (the reference file:line of each site is listed in DESIGN.md §RNG).
enum {
    XW_SITE_NAMES = 1, XW_SITE_MAZE = 2, XW_SITE_BLOCKS = 3, XW_SITE_GOAL_LOC = 4, XW_SITE_GOAL_ASSET = 5,
    XW_SITE_AGENT_LOC = 6, XW_SITE_TASK_A = 7, XW_SITE_TASK_B = 8, XW_SITE_TASK_SHUF = 9, XW_SITE_TASK_AGENT = 10,
    // 11 = XW_SITE_SENTENCE (xw_sentence.hpp)
    XW_SITE_AGENT_YAW = 12,  // set_property, --visible_radius > 0: agent yaw = choice(range(-1, 3)) * PI_2 (xworld_env.py:208-210)
    XW_SITE_GOAL_POSE = 13   // goal yaw / scale / offset = uniform(...) (xworld_env.py:211-223): index 4 * goal + {0, 1, 2}
};
// Goal yaws are drawn on a grid of XW_YAW_STEPS points of [0, 4 * PI_2): cos / sin of the rotation angle feed lrint() in
// cv::warpAffine's fixed-point set-up, and device libm does not agree with the host's to the last bit -- the host evaluates the
// 4096 possible angles once per handle (XwDev::yaw_cs) and the device only multiplies and adds (IEEE-exact).
#define XW_YAW_STEPS 4096

// directions relative to a heading (XWorld3DNavTargetDirection.__compute_triple_direction)
enum { XW_DIR_FALSE = 0, XW_DIR_FRONT = 1, XW_DIR_BEHIND = 2, XW_DIR_LEFT = 3, XW_DIR_RIGHT = 4 };

// XWorldNav._configure (XWorldNav.py:24-35): level l = a (3+l)-sided world inside the 8x8 map, padded with bricks
#define XW_N_T3 5
#define XW_WIN_SIZE 200
#define XW_WIN_WORDS 7
#define XW_N_LEVELS 6
XW_HD void xw_level_dims(int level, int& D, int& n_goals, int& n_blocks) {
    D = 3 + level;
    n_goals = level < 3 ? 2 : 4;                                   // num_goals_seq  = [2, 2, 2, 4, 4, 4]
    n_blocks = level < 5 ? 3 * level : 16;                          // num_blocks_seq = [0, 3, 6, 9, 12, 16]
}

struct XwDev {
    int32_t n;                 // envs on this device
    int32_t H, W, CS;          // map size, grid row stride (bytes)
    int32_t G, n_blocks, rules, max_steps, max_steps_factor, auto_reset;
    int32_t retry_width;       // attempts the warp of a queued env evaluates side by side in its first retry round
    int32_t vr;                // --visible_radius after the clamp to the map side; 0 = fully observed
    int32_t task_mode;         // xw_task_mode
    int32_t* n_invalid;        // [1] envs that were handed an invalid action since the last xw_step* call returned (or NULL)
    uint64_t seed;
    int64_t gid0;              // global id of env 0
    // ---- per-env state (SoA) ----
    uint8_t* grid;             // [n][CS]
    uint8_t *agent_x, *agent_y, *facing, *task, *stage, *event, *succ, *tmask, *aux0, *aux1, *aux2;
    uint8_t *goal_x, *goal_y;  // [XW_MAX_GOALS][n]
    int32_t* goal_icon;        // [XW_MAX_GOALS][n]
    int32_t* goal_name;        // [XW_MAX_GOALS][n]
    // ---- first-person view (vr > 0): Entity.yaw / scale / offset of the goals (xworld_env.py:211-223); NULL otherwise ----
    uint16_t* goal_yaw;        // [XW_MAX_GOALS][n] index on the yaw grid: yaw = 4 * PI_2 * idx / XW_YAW_STEPS
    double *goal_scale, *goal_offset;  // [XW_MAX_GOALS][n]
    const double* yaw_cs;      // [XW_YAW_STEPS][2] cos, sin of (90 - yaw * 180 / pi) degrees, evaluated by the host's libm
    int32_t *steps_in_task, *num_steps, *episode, *n_success, *n_failure, *success_steps, *error;
    uint32_t* minstd;
    unsigned long long* task_perf;  // [XW_N_T3][3] successes, failures, success steps per task class, summed over the batch
                                    // (Task::obtain_performance -> Teacher::report_task_performance, teacher.cpp:175-200)
    float* stage_rew;          // [n] the handle's own reward / game_over staging arrays (the host-buffer entry points), or NULL:
    int32_t* stage_over;       //     a reset clears the env's slots, so that an env that then sits a step out reads "alive", 0
    uint8_t* ctx_flag;         // [n] --context > 1 only: 1 = the env was stepped since its last render, 2 = it was reset
    // ---- curriculum (XWorldNav._configure with --curriculum > 0; all NULL / 0 when it is off) ----
    double curriculum;         // FLAGS_curriculum: (double)(float) of the option, py_simulator.cpp:127
    int32_t check_period;      // XWorldEnv.curriculum_check_period (xworld_env.py:58)
    uint8_t* level;            // [n] XWorldEnv.current_level
    int32_t* check_counter;    // [n] XWorldEnv.curriculum_check_counter
    uint8_t *win_len, *win_pos, *win_sum;  // [n][XW_N_T3]: XWorld3DTask.success_seq of each task class as a
    uint32_t* win_bits;        // [n][XW_N_T3][XW_WIN_WORDS] 200-entry ring of bits (xworld3d_task.py:47,129-133)
    // ---- catalog ----
    int32_t n_names, brick_icon, agent_icon;
    const int32_t *name_first, *name_icons;
    const uint8_t* icon_colored;
    // ---- auto-reset queue (ping-pong counters) ----
    int32_t* reset_count;      // [4]: [0..1] queue lengths, [2..3] positions claimed by the reset launch's warps (same ping-pong)
    int32_t* reset_list;       // [n]
};

// ------------------------------------------------------------------------------------ RNG
XW_HD uint32_t xw_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// Philox4x32-10 (Salmon et al. 2011); counter = (env id lo, hi, episode, attempt|site|index/4).
struct XwDraw4 { uint32_t v0, v1, v2, v3; };
XW_HD XwDraw4 xw_draw_block(uint64_t seed, int64_t gid, uint32_t episode, uint32_t attempt, uint32_t site, uint32_t block) {
    uint32_t c0 = (uint32_t)(uint64_t)gid, c1 = (uint32_t)((uint64_t)gid >> 32), c2 = episode;
    uint32_t c3 = ((attempt & 0xffu) << 24) | ((site & 0xffu) << 16) | (block & 0xffffu);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = xw_mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = xw_mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    XwDraw4 o; o.v0 = c0; o.v1 = c1; o.v2 = c2; o.v3 = c3;
    return o;
}
// A run of consecutive indices of one substream: one Philox evaluation per four draws instead of one per draw.
struct XwDrawSeq {
    uint64_t seed; int64_t gid; uint32_t episode, attempt, site, block; XwDraw4 cur;
};
XW_HD XwDrawSeq xw_draw_seq(uint64_t seed, int64_t gid, uint32_t episode, uint32_t attempt, uint32_t site) {
    XwDrawSeq q; q.seed = seed; q.gid = gid; q.episode = episode; q.attempt = attempt; q.site = site; q.block = 0xffffffffu;
    q.cur.v0 = q.cur.v1 = q.cur.v2 = q.cur.v3 = 0;
    return q;
}
XW_HD uint32_t xw_draw_next(XwDrawSeq& q, uint32_t index) {  // == xw_draw(..., index)
    if ((index >> 2) != q.block) { q.block = index >> 2; q.cur = xw_draw_block(q.seed, q.gid, q.episode, q.attempt, q.site, q.block); }
    const uint32_t lane = index & 3u;
    return lane == 0 ? q.cur.v0 : lane == 1 ? q.cur.v1 : lane == 2 ? q.cur.v2 : q.cur.v3;
}
XW_HD uint32_t xw_draw(uint64_t seed, int64_t gid, uint32_t episode, uint32_t attempt, uint32_t site, uint32_t index) {
    uint32_t c0 = (uint32_t)(uint64_t)gid, c1 = (uint32_t)((uint64_t)gid >> 32), c2 = episode;
    uint32_t c3 = ((attempt & 0xffu) << 24) | ((site & 0xffu) << 16) | ((index >> 2) & 0xffffu);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = xw_mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = xw_mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    uint32_t lane = index & 3u;
    return lane == 0 ? c0 : lane == 1 ? c1 : lane == 2 ? c2 : c3;
}
XW_HD uint32_t xw_randbelow(uint32_t u, uint32_t n) { return xw_mulhi32(u, n); }

// std::minstd_rand0 step + std::uniform_int_distribution<int>(0, size-1) as libstdc++ evaluates
// util::get_rand_ind (simulator_util.cpp:66-73) -- the teacher's task sampling stream.
XW_HD uint32_t xw_minstd_next(uint32_t& s) {
    s = (uint32_t)(((uint64_t)s * 16807ull) % 2147483647ull);
    return s;
}
XW_HD int32_t xw_get_rand_ind(uint32_t& s, int32_t size) {
    const uint32_t urngrange = 2147483645u;
    const uint32_t scaling = urngrange / (uint32_t)size;
    const uint32_t past = (uint32_t)size * scaling;
    uint32_t ret;
    do ret = xw_minstd_next(s) - 1u; while (ret >= past);
    return (int32_t)(ret / scaling);
}
