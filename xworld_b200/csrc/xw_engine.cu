// xw_engine.cu -- the C ABI of include/xworld_b200.h: handle management, kernel launches.
// Built in-tree as xworld_b200/libxworld_b200.so for sm_100a.  No torch types, no CPU fallback for
// the CUDA games: every entry point either launches kernels or returns an error.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <time.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "xw_common.cuh"
#include "xw_race.cuh"
#include "xw_wire.hpp"
#include "xw_render.cuh"
#include "xw_render_host.hpp"
#include "xw_sentence.hpp"
#include "xw_reset.cuh"
#include "xw_step.cuh"
#include "xw_fpv.cuh"
#include "xw_fpv_host.hpp"
#include "xw_teacher_names.hpp"

#include <cmath>
#include <functional>

// ------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(x)                                                                              \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if (e_ != cudaSuccess) return set_err(XW_ERR_CUDA, "%s: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------ kernels
// mode: list != NULL -> envs list[0..*count); else every env (optionally filtered by mask).
// List mode (the per-step auto-reset queue: a few per cent of the envs) gives every env a WARP of its own: a
// reset is one long data-dependent instruction stream (maze DFS, rejection loops), and 32 different ones in one
// warp serialise -- measured 490 us per step at C2 with one lane per env.  The lanes of the warp evaluate the
// episode's attempts side by side (xw_reset_env_warp).
__global__ void __launch_bounds__(128) k_reset(XwDev d, const uint8_t* mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n || (mask && !mask[i])) return;
    xw_reset_env(d, i);
    // an explicit reset (not the auto-reset of a step, whose game_over code is still to be read): the env is "alive", reward 0
    if (d.stage_over) { d.stage_over[i] = 0; d.stage_rew[i] = 0.f; }
}
// (`count` = reset_count + parity; the claim counter of the same parity sits two ints behind it.  Warps CLAIM queue positions
//  instead of striding over them: beside the painter only the CTAs on the reserved SMs are resident, and with a static stride the
//  positions of the CTAs that wait for an SM would wait with them -- with claims the resident warps drain a normal queue, and a
//  burst (every env of a batch that was started in phase timing out in the same step) is finished by the whole grid once the
//  painter has left, instead of by the dozen warps of the reserved SMs: 200 ms per such step before, profiles/r02_summary.md.)
//  max_claims > 0 bounds what one warp takes: the launch that runs beside the painter on a second stream must end with the
//  painter, a later launch on every SM sweeps up the rest of a burst.)
// (168 registers, 3 CTAs per SM.  Register caps were tried and are not better: -DXW_RESET_MINB=4 (128 registers, 4 CTAs per SM) C3
//  195.1 M against 196-198 M env-steps/s, =6 (80 registers, spills) 167.7 M -- a reset is one long dependent chain per warp.)
#ifdef XW_RESET_MINB
#define XW_RESET_BOUNDS __launch_bounds__(128, XW_RESET_MINB)
#else
#define XW_RESET_BOUNDS __launch_bounds__(128)
#endif
__global__ void XW_RESET_BOUNDS k_reset_list(XwDev d, const int32_t* list, const int32_t* count, int max_claims) {
    __shared__ uint32_t s_stack[4][XW_RESET_STACK_WORDS];
    __shared__ uint32_t s_draws[4][XW_RESET_DRAW_WORDS];
    const int cnt = *count, wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int32_t* claim = const_cast<int32_t*>(count) + 2;
    for (int taken = 0; max_claims <= 0 || taken < max_claims; ++taken) {
        int w = 0;
        if (lane == 0) w = atomicAdd(claim, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= cnt) break;
        xw_reset_env_warp(d, list[w], s_stack[wi], s_draws[wi]);
    }
}

__global__ void __launch_bounds__(256) k_step(XwDev d, const int32_t* __restrict__ actions, int act_rep,
                                              float* __restrict__ reward, int32_t* __restrict__ over, int parity) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0) { d.reset_count[parity ^ 1] = 0; d.reset_count[2 + (parity ^ 1)] = 0; }  // consumed by the previous step's reset launch
    if (e >= d.n) return;
    const int32_t a = actions[e];
    if (a == XW_ACTION_NONE) return;  // this env sits the step out: state and its reward / game_over slots untouched
    float r; int32_t o;
    bool need = xw_step_env(d, e, a, act_rep, &r, &o);
    reward[e] = r; over[e] = o;
    if (need) {  // warp-aggregated append to the auto-reset queue
        unsigned m = __activemask();
        int leader = __ffs(m) - 1, lane = threadIdx.x & 31;
        int base = 0;
        if (lane == leader) base = atomicAdd(&d.reset_count[parity], __popc(m));
        base = __shfl_sync(m, base, leader);
        d.reset_list[base + __popc(m & ((1u << lane) - 1))] = e;
    }
}

__global__ void __launch_bounds__(256) k_race_reset(XwRaceCfg r, const uint8_t* mask) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r.n || (mask && !mask[e])) return;
    xw_race_reset_env(r, e);
}
// 46 registers (5 CTAs / SM).  Forcing 6 or 8 CTAs per SM (40 / 32 registers, spills) is slower: 18.4 / 21.1 us
// against 18.2 us per step at 1,048,576 envs (profiles/r01_summary.md).
__global__ void __launch_bounds__(256) k_race_step(XwRaceCfg r, const int32_t* __restrict__ actions, int n_actions, int act_rep,
                                                   float* __restrict__ reward, int32_t* __restrict__ over, int32_t* error) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r.n) return;
    int a = actions[e];
    if (a == XW_ACTION_NONE) return;
    if (a < 0 || a >= n_actions) { error[e] = XW_ERR_INVALID_ACTION; reward[e] = 0.f; over[e] = 0; return; }
    float rw; int32_t o;
    XwRaceCar c = xw_race_load(r, e);
    const bool need = xw_race_step_car(r, c, a, act_rep, &rw, &o, r.state + (size_t)e * 4);
    reward[e] = rw; over[e] = o;
    if (need) xw_race_reset_car(r, c);  // the race reset is a few stores (four draws with --random): done in place
    xw_race_store(r, e, c);
}
// K consecutive take_actions calls of every env in ONE launch (xw_step_seq): actions / reward / game_over are [K][n], the car
// stays in registers between the steps, the state vector is written once (after the last step).  For open-loop action sequences
// (evaluation roll-outs, action-repeat style agents): a step of this game is ~60 bytes and ~300 instructions per env, so at one
// launch per step the launch itself is most of the time (profiles/r01_summary.md).
__global__ void __launch_bounds__(256) k_race_step_seq(XwRaceCfg r, const int32_t* __restrict__ actions, int n_actions, int act_rep, int K,
                                                       float* __restrict__ reward, int32_t* __restrict__ over, int32_t* error) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r.n) return;
    XwRaceCar c = xw_race_load(r, e);
    float st[4] = {r.state[(size_t)e * 4], r.state[(size_t)e * 4 + 1], r.state[(size_t)e * 4 + 2], r.state[(size_t)e * 4 + 3]};
    for (int k = 0; k < K; ++k) {
        const size_t i = (size_t)k * r.n + e;
        const int a = actions[i];
        if (a == XW_ACTION_NONE) continue;
        if (a < 0 || a >= n_actions) { error[e] = XW_ERR_INVALID_ACTION; reward[i] = 0.f; over[i] = 0; continue; }
        float rw; int32_t o;
        const bool need = xw_race_step_car(r, c, a, act_rep, &rw, &o, st);
        reward[i] = rw; over[i] = o;
        if (need) xw_race_reset_car(r, c);
    }
    xw_race_store(r, e, c);
#pragma unroll
    for (int q = 0; q < 4; ++q) r.state[(size_t)e * 4 + q] = st[q];
}

// ------------------------------------------------------------------------------------ handle
struct SimpleGameEnv {  // games/simple_game/simple_game_simulator.cpp (host; BASELINE config 1)
    int cur_pos; int64_t num_steps; std::vector<uint8_t> state; std::vector<float> rewards;
};

struct xw_sim {
    xw_config cfg;
    int n = 0, device = 0;
    cudaStream_t own_stream = nullptr, copy_stream = nullptr;
    // auto-reset beside the painter (step_xworld): the reset launch runs on its own stream on `reserve` SMs the painter leaves free
    cudaStream_t reset_stream = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    int32_t* h_reset_cnt = nullptr;   // pinned [2]: the queue lengths of the last steps (read back without a wait: sizes `reserve`)
    float reset_avg = 128.f;
    int reset_ctas_per_sm = 0;
    bool overlap_reset = false;
    bool reset_pdl = true;   // fully observed view: the reset kernel as the painter's programmatic dependent (same stream) instead of a second stream
    // XW_TRACE=1 (diagnostics): timestamps of one overlapped step, printed by xw_sync: k_step start / end, painter end, reset
    // end, re-paint end -- relative to the step's start
    bool trace = false;
    cudaEvent_t tr[6] = {};
    int trace_steps = 0;
    double host_enq_us = 0, host_sync_us = 0;   // step_hd: host time spent queueing the step / waiting for it (trace)
    cudaEvent_t tr_h0 = nullptr, tr_h1 = nullptr;   // trace: first / last operation of a synchronous step_hd on the handle's stream
    double gpu_span_us = 0, h2d_us = 0;
    int host_n = 0;
    cudaEvent_t ev_step = nullptr, ev_copy = nullptr, ev_frames = nullptr, ev_h2d = nullptr;
    int64_t launches = 0;
    int step_parity = 0;
    // xworld
    XwDev d;
    XwRender r;
    XwRenderTables tab;
    std::vector<void*> allocs;
    int n_sms = 148, render_grid = 0, render_smem = 0;
    bool render_sb = false, render_sp = false;
    uint8_t* tables = nullptr;      // phase atlas + edge / pair tables, one allocation
    size_t tables_bytes = 0, l2_window = 0;
    float l2_ratio = 0.f;
    void (*render_fn)(XwDev, XwRender, uint8_t*, size_t) = nullptr;  // k_render<WR> for this frame width
    void (*render_list_fn)(XwDev, XwRender, uint8_t*, size_t) = nullptr;  // the painter's list mode (same instantiation, LIST = true)
    int C = 3;                      // frame channels: 3 (planes B, G, R) or 1 (--color=false)
    // first-person view
    XwFpv fpv;
    bool fpv_fast = false;
    void (*fpv_cells_fn)(XwDev, XwFpv, uint8_t*, size_t, const int32_t*, const int32_t*, int, int, int, int, int, int, int) = nullptr;
    void (*fpv_stream_fn)(XwFpv, uint8_t*, size_t, int, int, int, int, int) = nullptr;
    bool fpv_streaming = false;  // XW_FPV_STREAM=1: the goal kernel BESIDE the frame kernel (k_fpv_goal_stream) -- measured slower than one
                                 // after the other (0.45 ms against 0.38 ms per render at 65,536 envs, profiles/r02_summary.md): opt-in
    int fpv_seq = 0;
    void (*fpv_goal_fn)(XwFpv, uint8_t*, size_t, int, int) = nullptr;
    int fpv_smem = 0, fpv_grid = 0, fpv_nt = 256, fpv_ng = 1, fpv_goal_grid = 0, fpv_goal_nt = 160;
    // the batch is rendered in fpv_chunks chunks: k_fpv_goal_cells of chunk k runs on fpv_stream beside the frame kernel of chunk k + 1
    enum { FPV_MAX_CHUNKS = 8 };
    int fpv_chunks = 1, fpv_parity = 0;
    cudaStream_t fpv_stream = nullptr;
    cudaEvent_t fpv_ev[FPV_MAX_CHUNKS + 1] = {};
    uint8_t* d_bgr = nullptr;       // --color=false: the colour frames the gray pass reads
    // race
    XwRaceCfg race;
    int32_t* race_error = nullptr;
    // simple game
    std::vector<SimpleGameEnv> sg;
    // host staging (pinned)
    int32_t *h_act = nullptr, *h_over = nullptr, *d_act = nullptr, *d_over = nullptr;
    int32_t* h_invalid = nullptr;   // pinned: running count of invalid actions (XwDev::n_invalid), read back by the host-buffer calls
    int32_t invalid_seen = 0;
    float *h_rew = nullptr, *d_rew = nullptr;
    uint8_t *d_mask = nullptr, *d_frames = nullptr;
    // timing
    bool timing = false;
    std::vector<cudaEvent_t> ev;      // pairs around the render launches
    size_t ev_used = 0;
    std::vector<cudaEvent_t> ev2;     // pairs around the step + reset (+ goal warp) launches
    size_t ev2_used = 0;
};

// Every entry point that takes a handle runs on the handle's device and leaves the caller's current device as it found it
// (two handles on different GPUs in one process, or a caller that switches devices between calls).
struct DevGuard {
    int prev = -1;
    explicit DevGuard(const xw_sim* s) {
        if (!s || s->cfg.game == XW_GAME_SIMPLE_GAME) return;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != s->device) { prev = cur; cudaSetDevice(s->device); }
    }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T>
static int dalloc(xw_sim* s, T** p, size_t count, bool zero = true) {
    void* q = nullptr;
    CUDA_TRY(cudaMalloc(&q, count * sizeof(T) + 16));
    if (zero) CUDA_TRY(cudaMemset(q, 0, count * sizeof(T) + 16));
    s->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}
template <typename T>
static int dupload(xw_sim* s, const T** p, const T* host, size_t count) {
    T* q = nullptr;
    int rc = dalloc(s, &q, count, false);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(q, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *p = q;
    return 0;
}

// `stream` is the caller's cudaStream_t; NULL is the legacy default stream, as in the CUDA runtime.
static cudaStream_t pick_stream(xw_sim*, void* stream) { return (cudaStream_t)stream; }

// ThreadCounter seed of the reference's i-th simulator thread (simulator_util.cpp:38-52):
// int seed = std::hash<std::string>()(std::to_string(FLAGS_simulator_seed + i)); reng_.seed(seed)
static uint32_t minstd_seed(int32_t simulator_seed, int64_t thread_no) {
    int32_t seed = (int32_t)(uint32_t)std::hash<std::string>()(std::to_string((int)(simulator_seed + thread_no)));
    uint64_t x = ((uint64_t)(int64_t)seed) % 2147483647ull;  // linear_congruential_engine::seed
    return x == 0 ? 1u : (uint32_t)x;
}

extern "C" {

const char* xw_last_error(void) { return g_err; }

void xw_config_init(xw_config* c) {
    memset(c, 0, sizeof *c);
    c->abi_version = XW_ABI_VERSION;
    c->game = XW_GAME_XWORLD;
    c->height = c->width = 8;  // XWorldNav.py:10-11
    c->n_goals = 4; c->n_blocks = 16;  // XWorldNav.py:31-32, last level
    c->rules = XW_RULES_NAV3D;
    c->context = 1;
    c->max_steps_factor = 10;
    c->array_size = 6;
    c->track_width = 20.f; c->track_length = 100.f; c->track_radius = 30.f;
    c->reward_scale = 1.f;
}

// First-person view: tables + the per-env goal icon cache (xw_fpv.cuh).
static int create_fpv(xw_sim* s, const xw_catalog* cat, int OH, int OW) {
    const xw_config& c = s->cfg;
    XwDev& d = s->d;
    XwFpv& F = s->fpv;
    memset(&F, 0, sizeof F);
    F.vr = d.vr; F.N = d.vr * 64; F.CH = c.height * 64; F.OH = OH; F.OW = OW; F.FB = 3 * OH * OW;
    F.ident1 = F.N == F.CH; F.ident2 = F.CH == OH && F.CH == OW;
    F.G = c.n_goals; F.brick_icon = cat->brick_icon; F.agent_icon = cat->agent_icon;
    XwRender& r = s->r;  // (dimensions only: xw_screen_dims and the context shift read them)
    memset(&r, 0, sizeof r);
    r.OH = OH; r.OW = OW; r.FB = F.FB; r.H = c.height; r.W = c.width; r.n_icons = cat->n_icons;
    int rc = 0;
    std::vector<int16_t> o, a0, a1;
    xw_fpv_resize_tables(F.N, F.CH, false, o, a0, a1);
    rc |= dupload(s, &F.x1ofs, o.data(), o.size()); rc |= dupload(s, &F.x1a0, a0.data(), a0.size()); rc |= dupload(s, &F.x1a1, a1.data(), a1.size());
    xw_fpv_resize_tables(F.N, F.CH, true, o, a0, a1);
    rc |= dupload(s, &F.y1ofs, o.data(), o.size()); rc |= dupload(s, &F.y1a0, a0.data(), a0.size()); rc |= dupload(s, &F.y1a1, a1.data(), a1.size());
    xw_fpv_resize_tables(F.CH, OW, false, o, a0, a1);
    rc |= dupload(s, &F.x2ofs, o.data(), o.size()); rc |= dupload(s, &F.x2a0, a0.data(), a0.size()); rc |= dupload(s, &F.x2a1, a1.data(), a1.size());
    xw_fpv_resize_tables(F.CH, OH, true, o, a0, a1);
    rc |= dupload(s, &F.y2ofs, o.data(), o.size()); rc |= dupload(s, &F.y2a0, a0.data(), a0.size()); rc |= dupload(s, &F.y2a1, a1.data(), a1.size());
    rc |= dupload(s, &F.atlas64, cat->atlas64, (size_t)cat->n_icons * 12288);
    {
        std::vector<uint8_t> a4 = xw_fpv_agent_icons(cat->atlas64 + (size_t)cat->agent_icon * 12288);
        rc |= dupload(s, &F.agent4, a4.data(), a4.size());
        std::vector<int16_t> itab(32 * 32 * 4);
        xw_fpv_build_itab(itab.data());
        rc |= dupload(s, &F.itab, itab.data(), itab.size());
        std::vector<double> cs = xw_fpv_yaw_table();
        rc |= dupload(s, &d.yaw_cs, cs.data(), cs.size());
    }
    rc |= dalloc(s, &F.gcache, (size_t)s->n * F.G * 4096, false);
    uint8_t *pmap = nullptr, *Tb = nullptr, *Ta = nullptr;
    rc |= dalloc(s, &pmap, (size_t)4 * OH * OW);
    rc |= dalloc(s, &Tb, (size_t)4 * 3 * OH * OW);
    rc |= dalloc(s, &Ta, (size_t)4 * 3 * OH * OW);
    if (rc) return rc;
    F.pmap = pmap; F.Tb = Tb; F.Ta = Ta;
    k_fpv_build_tables<<<s->n_sms * 4, 256, 0, s->own_stream>>>(F, pmap, Tb, Ta);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    s->fpv_fast = OW % 4 == 0 && F.FB % 16 == 0;
    if (s->fpv_fast && OW == OH && OW % F.vr == 0 && (OW / F.vr) % 4 == 0) {
        // regular geometry?  every frame pixel inside one cell, every bs x bs block one cell (pmap, copied back)
        const int bs = OW / F.vr, ncell = F.vr * F.vr;
        std::vector<uint8_t> pm((size_t)4 * OH * OW);
        CUDA_TRY(cudaStreamSynchronize(s->own_stream));
        CUDA_TRY(cudaMemcpy(pm.data(), pmap, pm.size(), cudaMemcpyDeviceToHost));
        std::vector<uint8_t> c2b((size_t)4 * ncell, 0xff);
        bool regular = true;
        for (int f = 0; f < 4 && regular; ++f)
            for (int p = 0; p < OH * OW && regular; ++p) {
                const int y = p / OW, x = p % OW, blk = (y / bs) * F.vr + x / bs;
                const uint8_t cell = pm[(size_t)f * OH * OW + p];
                if (cell == 0xff || cell >= ncell) { regular = false; break; }
                uint8_t& slot = c2b[(size_t)f * ncell + cell];
                if (slot == 0xff) slot = (uint8_t)blk; else if (slot != blk) regular = false;
            }
        for (uint8_t v : c2b) if (v == 0xff) regular = false;
        if (regular) {
            uint4* taps = nullptr;
            rc |= dalloc(s, &taps, (size_t)4 * OH * OW * 4, false);
            rc |= dupload(s, &F.cell2block, c2b.data(), c2b.size());
            std::vector<uint8_t> b2c(c2b.size());
            for (int f = 0; f < 4; ++f)
                for (int cell = 0; cell < ncell; ++cell) b2c[(size_t)f * ncell + c2b[(size_t)f * ncell + cell]] = (uint8_t)cell;
            rc |= dupload(s, &F.block2cell, b2c.data(), b2c.size());
            if (rc) return rc;
            F.regular = 1; F.bs = bs; F.taps4 = taps;
            rc |= dalloc(s, &F.goal_count, 2 * (size_t)XW_FPV_SLOTS);
            rc |= dalloc(s, &F.goal_list, (size_t)s->n * F.G, false);
            if (rc) return rc;
            k_fpv_build_taps<<<s->n_sms * 4, 256, 0, s->own_stream>>>(F, taps);  // (F.taps4 == taps)
            s->launches++;
            CUDA_TRY(cudaGetLastError());
        }
    }
    if (s->fpv_fast && F.regular) {
        const int ncell = F.vr * F.vr;
        s->fpv_nt = 128;
        const int per_group = (F.FB + d.CS + ((ncell + 15) & ~15) + 16 + 128 + 127) & ~127;
        s->fpv_ng = 8;
        while (s->fpv_ng > 1 && s->fpv_ng * per_group > 200 * 1024) s->fpv_ng >>= 1;
        s->fpv_smem = s->fpv_ng * per_group;
        s->fpv_nt = 128 * s->fpv_ng;
#define XW_FPV_PICK(NG)                                                                                                                  \
        (F.bs == 12 && F.vr == 7 ? k_render_fpv_cells<128, NG, 12, 7> : F.bs == 28 && F.vr == 3 ? k_render_fpv_cells<128, NG, 28, 3>       \
         : F.bs == 84 && F.vr == 1 ? k_render_fpv_cells<128, NG, 84, 1> : k_render_fpv_cells<128, NG, 0, 0>)
        s->fpv_cells_fn = s->fpv_ng == 8 ? XW_FPV_PICK(8) : s->fpv_ng == 4 ? XW_FPV_PICK(4) : s->fpv_ng == 2 ? XW_FPV_PICK(2) : XW_FPV_PICK(1);
#undef XW_FPV_PICK
        s->fpv_goal_fn = F.bs == 12 && F.vr == 7 ? k_fpv_goal_cells<12, 7> : F.bs == 28 && F.vr == 3 ? k_fpv_goal_cells<28, 3>
                       : F.bs == 84 && F.vr == 1 ? k_fpv_goal_cells<84, 1> : k_fpv_goal_cells<0, 0>;
        s->fpv_goal_nt = F.bs * F.bs >= 160 ? 160 : ((F.bs * F.bs + 31) & ~31);
        s->fpv_stream_fn = F.bs == 12 && F.vr == 7 ? k_fpv_goal_stream<12, 7> : F.bs == 28 && F.vr == 3 ? k_fpv_goal_stream<28, 3>
                         : F.bs == 84 && F.vr == 1 ? k_fpv_goal_stream<84, 1> : k_fpv_goal_stream<0, 0>;
        CUDA_TRY(cudaFuncSetAttribute(s->fpv_stream_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 16384 + 128));
        // the largest shared-memory carve-out for both: the frame kernel alone would get the smallest configuration that holds its
        // 172 KB (196 KB), and the carve-out of an SM cannot change while a CTA is resident -- the streaming kernel's 49 KB would
        // then not fit beside it and it would only start when the frame kernel has left
        if (const char* ev = getenv("XW_FPV_STREAM")) s->fpv_streaming = atoi(ev) != 0;
        if (F.G > 6) s->fpv_streaming = false;   // (the frame kernel keeps six pending list positions per env)
        if (s->fpv_streaming) {
            CUDA_TRY(cudaFuncSetAttribute(s->fpv_stream_fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            CUDA_TRY(cudaFuncSetAttribute(s->fpv_cells_fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        } else {
            CUDA_TRY(cudaFuncSetAttribute(s->fpv_cells_fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutDefault));
        }
        if (const char* ev = getenv("XW_FPV_STREAM")) s->fpv_streaming = atoi(ev) != 0;
        {
            int per_sm = 0;
            CUDA_TRY(cudaFuncSetAttribute(s->fpv_goal_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 16));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s->fpv_goal_fn, s->fpv_goal_nt, 16384 + 16));
            s->fpv_goal_grid = s->n_sms * (per_sm > 0 ? per_sm : 1);
        }
        CUDA_TRY(cudaFuncSetAttribute(s->fpv_cells_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, s->fpv_smem));
        if (const char* ev = getenv("XW_FPV_CHUNKS")) { const int v = atoi(ev); if (v >= 1 && v <= xw_sim::FPV_MAX_CHUNKS) s->fpv_chunks = v; }
        s->fpv_grid = s->n_sms;
    } else if (s->fpv_fast) {
        s->fpv_smem = F.FB + 2 * OH * OW + 256;
        int max_optin = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
        if (s->fpv_smem > max_optin) s->fpv_fast = false;
        else {
            CUDA_TRY(cudaFuncSetAttribute(k_render_fpv<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, s->fpv_smem));
            int per_sm = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_render_fpv<256>, 256, s->fpv_smem));
            s->fpv_grid = s->n_sms * (per_sm > 0 ? per_sm : 1);
        }
    }
    CUDA_TRY(cudaStreamSynchronize(s->own_stream));
    return 0;
}

static int create_xworld(xw_sim* s, const xw_catalog* cat) {
    const xw_config& c = s->cfg;
    const int n = s->n;
    if (!cat || !cat->atlas64 || cat->n_icons <= 0) return set_err(XW_ERR_INVALID_ARG, "xworld needs an icon catalog");
    if (c.height != c.width) return set_err(XW_ERR_INVALID_ARG, "only square maps (maze2d.py:78)");
    if (c.height < 3 || c.height > XW_MAX_DIM) return set_err(XW_ERR_INVALID_ARG, "map side must be in [3,%d]", XW_MAX_DIM);
    if (c.n_goals < 1 || c.n_goals > XW_MAX_GOALS) return set_err(XW_ERR_INVALID_ARG, "n_goals must be in [1,%d]", XW_MAX_GOALS);
    if (cat->n_names < c.n_goals) return set_err(XW_ERR_INVALID_ARG, "catalog has fewer goal names than n_goals");
    if (c.visible_radius < 0) return set_err(XW_ERR_INVALID_ARG, "visible_radius must be >= 0");
    const int vr = c.visible_radius < c.height ? c.visible_radius : c.height;  // xworld_simulator.cpp:63-64
    if (vr > 0 && vr % 2 == 0) return set_err(XW_ERR_INVALID_ARG, "visible_radius %d: must be an odd int (xmap.cpp:277)", vr);
    if (c.task_mode != XW_TASK_LANG_ACQUISITION && c.task_mode != XW_TASK_ONE_CHANNEL) return set_err(XW_ERR_INVALID_ARG, "unknown task_mode");
    if (c.task_mode == XW_TASK_ONE_CHANNEL && c.rules == XW_RULES_NAV2D && c.curriculum != 0)
        return set_err(XW_ERR_UNSUPPORTED, "one_channel + walls.json rules + curriculum: the task classes' usage records then depend on the "
                                           "XWorldRec question tasks, which are outside this engine");
    if (c.context < 1 || c.context > 16) return set_err(XW_ERR_INVALID_ARG, "context must be in [1,16]");
    if (c.rules != XW_RULES_NAV3D && c.rules != XW_RULES_NAV2D) return set_err(XW_ERR_INVALID_ARG, "unknown rules");
    if (c.curriculum != 0) {  // XWorldNav's level schedule is written for its own 8x8 map (XWorldNav.py:10-11,27-33)
        if (!(c.curriculum > 0)) return set_err(XW_ERR_INVALID_ARG, "curriculum must be >= 0");
        if (c.height != 8 || c.n_goals != 4 || c.n_blocks != 16)
            return set_err(XW_ERR_INVALID_ARG, "curriculum > 0 needs XWorldNav's own map: 8x8, 4 goals, 16 blocks");
        if (c.start_level < 0 || c.start_level >= XW_N_LEVELS) return set_err(XW_ERR_INVALID_ARG, "start_level must be in [0,%d]", XW_N_LEVELS - 1);
        if (c.curriculum_check_period < 0) return set_err(XW_ERR_INVALID_ARG, "curriculum_check_period must be >= 0");
    }
    {  // the maze must offer n_blocks wall cells ("too many blocks for a valid maze", xworld_env.py:443)
        int D = c.height, X = (D % 2 == 0) ? D - 1 : D, nx = (X + 1) / 2;
        int walls = X * X - nx * nx - (nx * nx - 1) + ((D % 2 == 0) ? (X / 2) + (D / 2) : 0);
        if (c.n_blocks < 0 || c.n_blocks > walls) return set_err(XW_ERR_INVALID_ARG, "n_blocks %d > %d maze wall cells", c.n_blocks, walls);
        if (D * D - walls < c.n_goals + 1) return set_err(XW_ERR_INVALID_ARG, "map too small for goals + agent");
    }
    XwDev& d = s->d;
    memset(&d, 0, sizeof d);
    d.n = n; d.H = c.height; d.W = c.width; d.CS = (c.height * c.width + 15) & ~15;
    d.G = c.n_goals; d.n_blocks = c.n_blocks; d.rules = c.rules; d.max_steps = c.max_steps;
    d.max_steps_factor = c.max_steps_factor; d.auto_reset = c.auto_reset;
    d.seed = c.seed; d.gid0 = c.env_id_offset;
    d.vr = vr; d.task_mode = c.task_mode;
    s->C = c.gray ? 1 : 3;
    { const char* eo = getenv("XW_RESET_OVERLAP"); s->overlap_reset = !eo || atoi(eo) != 0; }
    { const char* et = getenv("XW_TRACE"); s->trace = et && atoi(et) != 0; }
    { const char* ep = getenv("XW_RESET_PDL"); s->reset_pdl = !ep || atoi(ep) != 0; }
    { const char* ew = getenv("XW_RESET_RETRY_WIDTH"); int w = ew ? atoi(ew) : 8; d.retry_width = w < 1 ? 1 : (w > 32 ? 32 : w); }
    int rc = 0;
    rc |= dalloc(s, &d.grid, (size_t)n * d.CS);
    uint8_t** u8s[] = {&d.agent_x, &d.agent_y, &d.facing, &d.task, &d.stage, &d.event, &d.succ, &d.tmask, &d.aux0, &d.aux1, &d.aux2};
    for (auto p : u8s) rc |= dalloc(s, p, (size_t)n);
    rc |= dalloc(s, &d.goal_x, (size_t)n * XW_MAX_GOALS);
    rc |= dalloc(s, &d.goal_y, (size_t)n * XW_MAX_GOALS);
    rc |= dalloc(s, &d.goal_icon, (size_t)n * XW_MAX_GOALS);
    rc |= dalloc(s, &d.goal_name, (size_t)n * XW_MAX_GOALS);
    int32_t** i32s[] = {&d.steps_in_task, &d.num_steps, &d.episode, &d.n_success, &d.n_failure, &d.success_steps, &d.error};
    for (auto p : i32s) rc |= dalloc(s, p, (size_t)n);
    rc |= dalloc(s, &d.minstd, (size_t)n);
    rc |= dalloc(s, &d.reset_count, 4);
    rc |= dalloc(s, &d.reset_list, (size_t)n);
    rc |= dalloc(s, &d.n_invalid, 1);
    rc |= dalloc(s, &d.task_perf, (size_t)XW_N_T3 * 3);
    if (c.context > 1) rc |= dalloc(s, &d.ctx_flag, (size_t)n);
    if (vr > 0) {
        rc |= dalloc(s, &d.goal_yaw, (size_t)n * XW_MAX_GOALS);
        rc |= dalloc(s, &d.goal_scale, (size_t)n * XW_MAX_GOALS);
        rc |= dalloc(s, &d.goal_offset, (size_t)n * XW_MAX_GOALS);
    }
    if (c.curriculum != 0) {
        d.curriculum = (double)c.curriculum;
        d.check_period = c.curriculum_check_period > 0 ? c.curriculum_check_period : 100;
        rc |= dalloc(s, &d.level, (size_t)n);
        rc |= dalloc(s, &d.check_counter, (size_t)n);
        rc |= dalloc(s, &d.win_len, (size_t)n * XW_N_T3);
        rc |= dalloc(s, &d.win_pos, (size_t)n * XW_N_T3);
        rc |= dalloc(s, &d.win_sum, (size_t)n * XW_N_T3);
        rc |= dalloc(s, &d.win_bits, (size_t)n * XW_N_T3 * XW_WIN_WORDS);
        if (!rc && c.start_level) CUDA_TRY(cudaMemset(d.level, c.start_level, (size_t)n));
    }
    if (rc) return rc;
    {
        std::vector<uint32_t> seeds(n);
        for (int i = 0; i < n; ++i) seeds[i] = minstd_seed(c.simulator_seed, c.env_id_offset + i + 1);
        CUDA_TRY(cudaMemcpy(d.minstd, seeds.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice));
    }
    // catalog
    d.n_names = cat->n_names; d.brick_icon = cat->brick_icon; d.agent_icon = cat->agent_icon;
    rc |= dupload(s, &d.name_first, cat->name_first, (size_t)cat->n_names + 1);
    rc |= dupload(s, &d.name_icons, cat->name_icons, (size_t)cat->name_first[cat->n_names]);
    rc |= dupload(s, &d.icon_colored, cat->icon_colored, (size_t)cat->n_icons);
    if (rc) return rc;
    // renderer
    int OH = c.out_h > 0 ? c.out_h : c.height * 12, OW = c.out_w > 0 ? c.out_w : c.width * 12;  // xworld_simulator.cpp:52-61
    if (vr > 0) {  // block_size = 84 / visible_radius, visible_radius blocks a side (xworld_simulator.cpp:62-68)
        if (c.out_h <= 0) OH = vr * (84 / vr);
        if (c.out_w <= 0) OW = vr * (84 / vr);
    }
    if (OH > XW_MAX_OUT || OW > XW_MAX_OUT) return set_err(XW_ERR_INVALID_ARG, "frame side must be <= %d", XW_MAX_OUT);
    if (c.gray) rc |= dalloc(s, &s->d_bgr, (size_t)n * 3 * OH * OW, false);
    if (rc) return rc;
    if (vr > 0) return create_fpv(s, cat, OH, OW);
    if (OH > c.height * 64 || OW > c.width * 64) return set_err(XW_ERR_UNSUPPORTED, "fully observed frames larger than the %d-px canvas (upscaling)", c.height * 64);
    s->tab = xw_build_render_tables(c.height, c.width, OH, OW);
    XwRenderTables& t = s->tab;
    XwRender& r = s->r;
    memset(&r, 0, sizeof r);
    r.OH = OH; r.OW = OW; r.WR = t.WR; r.FB = t.FB; r.H = c.height; r.W = c.width;
    r.n_sr = (int)t.sr.size();
    r.n_icons = cat->n_icons; r.brick_icon = cat->brick_icon; r.agent_icon = cat->agent_icon;
#if defined(XW_SP_DEBUG)
    { const char* ed = getenv("XW_RENDER_DEBUG"); r.debug = ed ? atoi(ed) : 0; }
#endif
#if defined(XW_SP_PROF)
    { unsigned int* pp = nullptr; rc |= dalloc(s, &pp, 16); r.prof = pp; }
#endif
    rc |= dupload(s, &r.taps.xofs, t.xofs.data(), t.xofs.size());
    rc |= dupload(s, &r.taps.xa0, t.xa0.data(), t.xa0.size());
    rc |= dupload(s, &r.taps.xa1, t.xa1.data(), t.xa1.size());
    rc |= dupload(s, &r.taps.yofs, t.yofs.data(), t.yofs.size());
    rc |= dupload(s, &r.taps.ya0, t.ya0.data(), t.ya0.size());
    rc |= dupload(s, &r.taps.ya1, t.ya1.data(), t.ya1.size());
    if (t.fast_ok) {
        // warp groups per CTA: as many pairs of private frame buffers as shared memory holds
        // (XW_RENDER_GROUPS / XW_RENDER_GROUP_THREADS / XW_RENDER_SPLIT_M3 override the choice, for tuning)
        int max_optin = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
        const char* eg = getenv("XW_RENDER_GROUPS");
        const char* et = getenv("XW_RENDER_GROUP_THREADS");
        const char* es = getenv("XW_RENDER_SPLIT_M3");
        const bool split = es ? atoi(es) != 0 : false;
        const char* ec = getenv("XW_RENDER_CONFLICT_FREE");
        const bool cfree = ec ? atoi(ec) != 0 : false;
        // "sb" (default): one frame buffer per group, up to 8 groups; "pipe": two per group, up to 4 groups.
        // Measured on B200 (profiles/r01_summary.md): of the two, sb 8x64 threads is the faster at 84x84 frames.
        const char* em = getenv("XW_RENDER_MODE");
        s->render_sb = !(em && !strcmp(em, "pipe"));
        // "sp" (default when the geometry allows): the sparse painter -- white pre-fill + the words of the
        // non-white cells only; "sb" / "pipe": the dense plan compositors
        xw_build_paint_tables(t);
        s->render_sp = t.sp_ok && !(em && (!strcmp(em, "pipe") || !strcmp(em, "sb")));
        if (s->render_sp) {
            r.nwc = t.nwc; r.ns = t.ns; r.n_sc = (int)t.sc.size();
            r.ctab_merge = xw_want_ctab_merge(r, t, max_optin) ? 1 : 0;
            if (r.ctab_merge) { r.ns = t.ns_strad; t.wshare = t.wshare_strad; t.ns = t.ns_strad; }
            r.slot_magic = 65536 / t.nwc + 1;
            const char* ef = getenv("XW_RENDER_SP_FILL");
            r.sp_fill = ef ? atoi(ef) : 0;
        }
        const int max_groups = s->render_sb ? XW_RENDER_MAX_GROUPS : XW_RENDER_MAX_GROUPS / 2;
        int G = eg ? atoi(eg) : max_groups;
        if (G > max_groups) G = max_groups;
        bool found = false;
        for (; G >= 1 && !found; --G) {
            // <= 512 threads/CTA keeps 124 registers per thread (no spills, 24-word load batches)
            // (k_render_sp needs its 128 registers: <= 512 threads per CTA)
            int GT = et ? atoi(et) : (G >= 6 ? 64 : ((s->render_sp ? 512 : 768) / G) / 32 * 32);
            // 9 groups fit at 84x84 but need > 512 threads = fewer registers per thread: measured 30 % slower
            if (!eg && !s->render_sp && s->render_sb && G * GT > 512 && G > 8) continue;
            GT = GT / 32 * 32;
            if (s->render_sp) {  // one special slot-plane per thread of a group, at least two warps (k_render_sp)
                const int need = 3 * (1 + c.n_goals) * t.nwc;
                if (GT < 64) GT = 64;
                if (GT < need) GT = (need + 31) / 32 * 32;
            }
            if (GT < 32 || G * GT > XW_RENDER_THREADS) { if (et) break; continue; }
            if (!eg && !et && s->render_sp && G * GT > 512) continue;  // the painter needs its 124 registers: no spills
            const char* e2 = getenv("XW_RENDER_TWO_PHASE");
            xw_build_plan(t, GT / 32, split, cfree, !s->render_sb, e2 && atoi(e2) != 0);
            r.n_plan = (int)t.plan.size(); r.n_plan1 = t.n_plan1;
            if ((s->render_sp ? xw_render_sp_smem(r, G).total : xw_render_smem(r, s->render_sb ? G : 2 * G).total) > max_optin) continue;
            r.G = G; r.GT = GT;
            found = true;
        }
        if (!found) t.fast_ok = false;
    }
    // All render tables live in ONE allocation so that a single L2 access-policy window can pin them:
    // a step streams 1.4 GB of frames through the 126 MB L2, which would otherwise evict the 8-20 MB of
    // tables and turn every special-cell look-up into a DRAM read (seen as 61 MB of DRAM reads per launch).
    {
        static const int16_t zero16 = 0;
        r.RB = t.fast_ok ? t.RB : 4;
        r.n_sc = (int)t.sc.size();
        const size_t n_T = (size_t)cat->n_icons * r.FB + XW_TABLE_PAD;
        const size_t n_ecol = ((size_t)(cat->n_icons + 1) * 2 * 3 * c.height * r.RB) * 2 + XW_TABLE_PAD;
        const size_t n_uv = ((size_t)(cat->n_icons + 1) * r.n_sr * 2 * 3 * OW + 8) * 2;
        const size_t n_corner = (size_t)(cat->n_icons + 1) * 3 * 4;
        const size_t n_col = (size_t)(cat->n_icons + 1) * 2 * r.n_sc * 3 * c.height * r.RB + 64, n_row = (size_t)(cat->n_icons + 1) * 2 * r.n_sr * 3 * OW + 64;
        const size_t n_cwb = (size_t)16 * r.n_sr * r.n_sc * 3 + 64;
        auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
        const size_t n_white = s->render_sp ? (size_t)r.FB + xw_ctab_words(r) * 4 + (size_t)(cat->n_icons + 1) * 16 : 0;
        if (s->render_sp) r.tc_rows = t.max_band_rows <= 8 ? 8 : 12;
        // cell-major copy of the phase atlas for the painter's special slots -- as long as it stays L2-resident next to
        // the other tables (C3: 12.6 MB; C4 would be 35 MB and measured slower than reading the atlas row by row)
        size_t n_tc = s->render_sp ? (size_t)cat->n_icons * c.height * c.width * 3 * t.nwc * r.tc_rows * 4 + 64 : 0;
        if (n_tc > ((size_t)24 << 20)) n_tc = 0;
        const size_t n_row2 = s->render_sp ? (size_t)(cat->n_icons + 1) * 2 * r.n_sr * r.WR * 3 * 4 + 64 : 0;
        const size_t total = up(n_T) + (t.fast_ok ? up(n_ecol) + up(n_uv) + up(n_corner) + 2 * up(n_col) + 2 * up(n_row) + up(n_cwb) + up(n_white) + up(n_tc) + 2 * up(n_row2) : 0);
        uint8_t* base = nullptr;
        rc |= dalloc(s, &base, total, false);
        if (rc) return rc;
        s->tables = base; s->tables_bytes = total;
        size_t o = 0;
        auto take = [&](size_t n) { uint8_t* p = base + o; o += up(n); return p; };
        r.T = take(n_T);
        if (t.fast_ok) {
            r.ecol = (const uint16_t*)take(n_ecol); r.uv = (const uint16_t*)take(n_uv); r.corner = (const uint32_t*)take(n_corner);
            r.colL = take(n_col); r.colR = take(n_col); r.rowT = take(n_row); r.rowB = take(n_row); r.cornerWB = take(n_cwb);
            if (s->render_sp) {
                uint8_t* w = take(n_white);
                CUDA_TRY(cudaMemset(w, 0xff, (size_t)r.FB));
                r.white = w;
                r.ctab = (const uint32_t*)(w + r.FB);  // (FB % 16 == 0)
                r.cornerP = r.ctab + xw_ctab_words(r);
                r.TC = n_tc ? (const uint32_t*)take(n_tc) : nullptr;
                r.rowT2 = (const uint32_t*)take(n_row2); r.rowB2 = (const uint32_t*)take(n_row2);
                rc |= dupload(s, &r.cellgeo, t.cellgeo.data(), t.cellgeo.size());
                rc |= dupload(s, &r.wcol, t.wcol.data(), t.wcol.size());
                rc |= dupload(s, &r.wshare, t.wshare.data(), t.wshare.size());
                static const uint8_t zero8 = 0;
                rc |= dupload(s, &r.sr_ty, t.sr_ty.empty() ? &zero8 : t.sr_ty.data(), t.sr_ty.empty() ? 1 : t.sr_ty.size());
            }
            rc |= dupload(s, &r.plan, t.plan.data(), t.plan.size());
            rc |= dupload(s, &r.cellinfo, t.cellinfo.data(), t.cellinfo.size());
            rc |= dupload(s, &r.sr, t.sr.empty() ? &zero16 : t.sr.data(), t.sr.empty() ? 1 : t.sr.size());
            rc |= dupload(s, &r.sc, t.sc.empty() ? &zero16 : t.sc.data(), t.sc.empty() ? 1 : t.sc.size());
            rc |= dupload(s, &r.band_y0, t.band_y0.data(), t.band_y0.size());
        }
        // L2 set-aside for persisting lines (best effort: not every driver / MIG slice grants it)
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, s->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, s->device);
        const char* ep = getenv("XW_RENDER_L2_PERSIST");
        if ((!ep || atoi(ep) != 0) && max_persist > 0 && max_window > 0) {
            size_t want = total + (total >> 2);
            if (want > (size_t)max_persist) want = (size_t)max_persist;
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
                s->l2_window = total < (size_t)max_window ? total : (size_t)max_window;
                s->l2_ratio = want >= s->l2_window ? 1.0f : (float)want / (float)s->l2_window;
            } else cudaGetLastError();
        }
    }
    rc |= dupload(s, &r.atlas64, cat->atlas64, (size_t)cat->n_icons * 64 * 64 * 3);
    if (rc) return rc;
    k_build_phase_atlas<<<s->n_sms * 8, 256, 0, s->own_stream>>>(r);
    s->launches++;
    if (t.fast_ok) {
        k_build_edge_tables<<<s->n_sms * 2, 256, 0, s->own_stream>>>(r);
        k_build_pair_tables<<<s->n_sms * 4, 256, 0, s->own_stream>>>(r);
        s->launches += 2;
        if (s->render_sp) {
            k_build_class_tables<<<s->n_sms * 2, 256, 0, s->own_stream>>>(r);
            if (r.TC) k_build_cell_tables<<<s->n_sms * 8, 256, 0, s->own_stream>>>(r);  // (after k_build_phase_atlas, same stream)
            k_build_row2_tables<<<s->n_sms, 256, 0, s->own_stream>>>(r);                  // (after k_build_pair_tables)
            s->launches += 3;
        }
    }
    CUDA_TRY(cudaGetLastError());
    if (t.fast_ok) {
        if (s->render_sp) r.sp = xw_render_sp_smem(r, r.G);
        s->render_smem = s->render_sp ? r.sp.total : xw_render_smem(r, s->render_sb ? r.G : 2 * r.G).total;
        // instantiations: compile-time row stride for the common frame widths x register budget by CTA size
        const int nt = r.G * r.GT;
#define XW_PICK(WR_) (nt <= 512 ? k_render<WR_, 512> : nt <= 768 ? k_render<WR_, 768> : k_render<WR_, 1024>)
#define XW_PICK_SB(WR_) (nt <= 512 ? k_render_sb<WR_, 512> : nt <= 576 ? k_render_sb<WR_, 576> : nt <= 768 ? k_render_sb<WR_, 768> : k_render_sb<WR_, 1024>)
        // (768 threads = 80 registers: only with the role-disjoint register sharing, i.e. 3 warps per group at 84x84)
        const bool disj = 3 * (1 + c.n_goals) * t.nwc + r.n_sr * r.WR <= r.GT;
#define XW_PICK_SP(WR_) (t.max_band_rows <= 8 ? (nt <= 512 ? k_render_sp<WR_, 512, 8> : nt <= 768 && disj ? k_render_sp<WR_, 768, 8, true> : k_render_sp<WR_, 1024, 8>) \
                                              : (nt <= 512 ? k_render_sp<WR_, 512, 12> : nt <= 768 && disj ? k_render_sp<WR_, 768, 12, true> : k_render_sp<WR_, 1024, 12>))
        // (the painter's compile-time row stride also fixes the frame height: square frames only)
#define XW_PICK_SPL(WR_) (t.max_band_rows <= 8 ? (nt <= 512 ? k_render_sp<WR_, 512, 8, false, true> : nt <= 768 && disj ? k_render_sp<WR_, 768, 8, true, true> : k_render_sp<WR_, 1024, 8, false, true>) \
                                               : (nt <= 512 ? k_render_sp<WR_, 512, 12, false, true> : nt <= 768 && disj ? k_render_sp<WR_, 768, 12, true, true> : k_render_sp<WR_, 1024, 12, false, true>))
        if (s->render_sp) {
            s->render_list_fn = r.OH != r.OW ? XW_PICK_SPL(0) : r.WR == 21 ? XW_PICK_SPL(21) : r.WR == 24 ? XW_PICK_SPL(24) : r.WR == 32 ? XW_PICK_SPL(32) : XW_PICK_SPL(0);
            CUDA_TRY(cudaFuncSetAttribute(s->render_list_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, s->render_smem));
        }
#undef XW_PICK_SPL
        if (s->render_sp) s->render_fn = r.OH != r.OW ? XW_PICK_SP(0) : r.WR == 21 ? XW_PICK_SP(21) : r.WR == 24 ? XW_PICK_SP(24) : r.WR == 32 ? XW_PICK_SP(32) : XW_PICK_SP(0);
        else if (s->render_sb) s->render_fn = r.WR == 21 ? XW_PICK_SB(21) : r.WR == 24 ? XW_PICK_SB(24) : r.WR == 32 ? XW_PICK_SB(32) : XW_PICK_SB(0);
        else s->render_fn = r.WR == 21 ? XW_PICK(21) : r.WR == 24 ? XW_PICK(24) : r.WR == 32 ? XW_PICK(32) : XW_PICK(0);
#undef XW_PICK
#undef XW_PICK_SB
#undef XW_PICK_SP
        CUDA_TRY(cudaFuncSetAttribute(s->render_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, s->render_smem));
        s->render_grid = s->n_sms;
    }
    CUDA_TRY(cudaStreamSynchronize(s->own_stream));
    return 0;
}

static int create_race(xw_sim* s) {
    const xw_config& c = s->cfg;
    XwRaceCfg& r = s->race;
    memset(&r, 0, sizeof r);
    r.n = s->n; r.track_type = c.track_type; r.full_manouver = c.race_full_manouver; r.difficulty = c.difficulty;
    r.max_steps = c.max_steps; r.auto_reset = c.auto_reset; r.reward_scale = (double)c.reward_scale;
    r.mid_x = 480 / 2; r.mid_y = 720 / 2;  // WINDOW_WIDTH/HEIGHT, simple_race_simulator.cpp:31-32,446
    if (c.track_type == 0) {  // StraightTrack ctor :103-109
        r.length = c.track_length; r.width = c.track_width;
        r.start_y = r.mid_y - (float)(0.4 * r.length);
        r.end_y = r.mid_y + (float)(0.6 * r.length);
        r.start_px = r.mid_x - 0.0f; r.start_py = r.start_y;
    } else {  // CircleTrack ctor :50-54, get_start_pos :76-79
        r.inner = c.track_radius; r.width = c.track_width; r.outer = r.inner + r.width;
        r.start_px = (r.inner + r.width / 2) + r.mid_x; r.start_py = 0.0f + r.mid_y;
    }
    int rc = 0;
    rc |= dalloc(s, &r.pos_x, (size_t)s->n);
    rc |= dalloc(s, &r.pos_y, (size_t)s->n);
    rc |= dalloc(s, &r.angle, (size_t)s->n);
    rc |= dalloc(s, &r.state, (size_t)s->n * 4);
    rc |= dalloc(s, &r.steps, (size_t)s->n);
    rc |= dalloc(s, &s->race_error, (size_t)s->n);
    r.random = c.race_random != 0;
    if (r.random && !rc) {  // env i = the reference's (i + 1)-th simulator thread (simulator_util.cpp:38-52)
        rc |= dalloc(s, &r.minstd, (size_t)s->n);
        if (rc) return rc;
        std::vector<uint32_t> seeds(s->n);
        for (int i = 0; i < s->n; ++i) seeds[i] = minstd_seed(c.simulator_seed, c.env_id_offset + i + 1);
        CUDA_TRY(cudaMemcpy(r.minstd, seeds.data(), sizeof(uint32_t) * s->n, cudaMemcpyHostToDevice));
    }
    return rc;
}

int xw_create(const xw_config* cfg, const xw_catalog* catalog, int32_t n_envs, int32_t device, xw_sim** out) {
    if (!cfg || !out) return set_err(XW_ERR_INVALID_ARG, "null argument");
    if (cfg->abi_version != XW_ABI_VERSION) return set_err(XW_ERR_INVALID_ARG, "abi_version %d != %d", cfg->abi_version, XW_ABI_VERSION);
    if (n_envs < 1) return set_err(XW_ERR_INVALID_ARG, "n_envs must be >= 1");
    xw_sim* s = new xw_sim();
    s->cfg = *cfg;
    s->n = n_envs;
    int rc = 0;
    if (cfg->game == XW_GAME_SIMPLE_GAME) {
        if (cfg->array_size < 2 || cfg->array_size > 4096) { delete s; return set_err(XW_ERR_INVALID_ARG, "array_size"); }
        s->sg.resize(n_envs);
        *out = s;
        xw_reset_host(s, nullptr, nullptr);
        return 0;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        delete s;
        return set_err(XW_ERR_NO_DEVICE, "no CUDA device: the xworld / simple_race paths have no CPU fallback");
    }
    int caller_device = -1;
    cudaGetDevice(&caller_device);
    if (device < 0) device = caller_device;
    s->device = device;
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{caller_device};  // the caller's current device is left as found
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s->n_sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) { delete s; return set_err(XW_ERR_CUDA, "device init: %s", cudaGetErrorString(e)); }
    if (cfg->game == XW_GAME_XWORLD) rc = create_xworld(s, catalog);
    else if (cfg->game == XW_GAME_SIMPLE_RACE) rc = create_race(s);
    else rc = set_err(XW_ERR_INVALID_ARG, "unknown game %d", cfg->game);
    if (rc) { xw_destroy(s); return rc; }
    *out = s;
    return 0;
}

void xw_destroy(xw_sim* s) {
    if (!s) return;
    DevGuard dev_guard(s);
    if (s->own_stream) cudaStreamSynchronize(s->own_stream);
#if defined(XW_SP_PROF)
    if (s->cfg.game == XW_GAME_XWORLD && s->r.prof && s->render_sp) {  // phase clocks summed over every render launch of the handle
        unsigned int h[16];
        cudaDeviceSynchronize();
        if (cudaMemcpy(h, s->r.prof, sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess) {
            const double envs = (double)((s->n + s->render_grid * s->r.G - 1) / (s->render_grid * s->r.G)) * (double)(s->launches > 8 ? s->launches - 4 : 1);
            fprintf(stderr, "XW_SP_PROF cycles/env (~%g envs): lane0 cells %.0f | barrier A %.0f | finish special %.0f | bricks %.0f | issue %.0f | barrier C %.0f | rest %.0f ;"
                            " lane63 drain %.0f | fill %.0f | barrier A %.0f | finish rword %.0f | bricks %.0f | issue %.0f | barrier C %.0f | rest %.0f\n",
                    envs, h[1] / envs, h[2] / envs, h[3] / envs, h[4] / envs, h[5] / envs, h[6] / envs, h[7] / envs,
                    h[8] / envs, h[9] / envs, h[10] / envs, h[11] / envs, h[12] / envs, h[13] / envs, h[14] / envs, h[15] / envs);
        }
    }
#endif
    for (void* p : s->allocs) cudaFree(p);
    if (s->h_act) cudaFreeHost(s->h_act);
    if (s->h_over) cudaFreeHost(s->h_over);
    if (s->h_rew) cudaFreeHost(s->h_rew);
    if (s->h_invalid) cudaFreeHost(s->h_invalid);
    if (s->h_reset_cnt) cudaFreeHost(s->h_reset_cnt);
    if (s->reset_stream) { cudaStreamSynchronize(s->reset_stream); cudaStreamDestroy(s->reset_stream); cudaEventDestroy(s->ev_a); cudaEventDestroy(s->ev_b); }
    for (auto ev : s->ev) cudaEventDestroy(ev);
    for (auto ev : s->ev2) cudaEventDestroy(ev);
    if (s->copy_stream) { cudaStreamDestroy(s->copy_stream); cudaEventDestroy(s->ev_step); cudaEventDestroy(s->ev_copy); cudaEventDestroy(s->ev_h2d); }
    if (s->ev_frames) cudaEventDestroy(s->ev_frames);
    if (s->fpv_stream) { cudaStreamSynchronize(s->fpv_stream); cudaStreamDestroy(s->fpv_stream); for (auto& e : s->fpv_ev) cudaEventDestroy(e); }
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
}

int32_t xw_num_envs(const xw_sim* s) { return s->n; }

int32_t xw_num_actions(const xw_sim* s) {
    if (s->cfg.game == XW_GAME_XWORLD) return s->d.vr == 0 ? 4 : 6;  // xitem.cpp:82-86
    if (s->cfg.game == XW_GAME_SIMPLE_GAME) return 2;
    return s->cfg.race_full_manouver ? 9 : 2;  // simple_race_simulator.cpp:432-440
}

int xw_screen_dims(const xw_sim* s, int32_t* h, int32_t* w, int32_t* c, int32_t* context) {
    if (s->cfg.game == XW_GAME_XWORLD) { *h = s->r.OH; *w = s->r.OW; *c = s->C; }
    else if (s->cfg.game == XW_GAME_SIMPLE_GAME) { *h = 1; *w = s->cfg.array_size; *c = 1; }  // simple_game_simulator.cpp:101-107
    else { *h = 1; *w = 4; *c = 1; }
    *context = s->cfg.context;
    return 0;
}

size_t xw_frame_bytes(const xw_sim* s) {
    int32_t h, w, c, k;
    xw_screen_dims(s, &h, &w, &c, &k);
    return (size_t)h * w * c * k * (s->cfg.game == XW_GAME_SIMPLE_RACE ? sizeof(float) : 1);
}

int xw_sentence_compose(const xw_sentence_query* q, char* buf, size_t cap) {
    if (!q || !buf || cap == 0) return set_err(XW_ERR_INVALID_ARG, "null buffer");
    if (q->rules != XW_RULES_NAV3D && q->rules != XW_RULES_NAV2D) return set_err(XW_ERR_INVALID_ARG, "unknown rules");
    const bool nav3d = q->rules == XW_RULES_NAV3D;
    if (q->task < 0 || q->task > (nav3d ? 4 : 3)) return set_err(XW_ERR_INVALID_ARG, "unknown task");
    using namespace xw_sentence;
    // (the tasks of walls.json that never leave idle in the reference commit have no command: DESIGN.md §4)
    if (!nav3d && q->kind == XW_SENT_START && (q->task == XW_T2_NEAR || q->task == XW_T2_BETWEEN)) { buf[0] = 0; return 0; }
    if (!nav3d && q->kind == XW_SENT_WRONG) { buf[0] = 0; return 0; }
    Grammar g = task_grammar(q->rules, q->task, quoted(q->name1), quoted(q->color), "'east'");
    if (nav3d && q->task == XW_T3_BETWEEN) g.add("G2", quoted(q->name2));
    if (!nav3d && q->task == XW_T2_BETWEEN) g.add("T", quoted(q->name2));
    const char* kinds3[4] = {"start", "correct", "wrong", "timeup"};
    const char* kinds2[4] = {"start", "finish", "finish", "timeup"};
    if (q->kind < 0 || q->kind > 3 || !g.bind("S", nav3d ? kinds3[q->kind] : kinds2[q->kind])) return set_err(XW_ERR_INVALID_ARG, "unknown sentence kind");
    if (q->kind == XW_SENT_START) {
        if (!q->name1 || !q->name1[0]) return set_err(XW_ERR_INVALID_ARG, "start sentence needs a goal name");
        if (nav3d && q->task == XW_T3_DIRECTION) {
            const char* dirs[5] = {nullptr, "FRONT", "BEHIND", "LEFT", "RIGHT"};
            if (q->direction < 1 || q->direction > 4 || !g.bind("P", dirs[q->direction])) return set_err(XW_ERR_INVALID_ARG, "direction must be 1..4");
        }
        if (((nav3d && q->task == XW_T3_BETWEEN)) && (!q->name2 || !q->name2[0])) return set_err(XW_ERR_INVALID_ARG, "Between needs two goal names");
        if (!nav3d && q->task == XW_T2_COLOR_TARGET && (!q->color || !q->color[0])) return set_err(XW_ERR_INVALID_ARG, "ColorTarget needs a colour");
    }
    uint32_t i = 0;
    auto draw = [&]() { return xw_draw(q->seed, q->env_id, q->episode, 0, XW_SITE_SENTENCE, ((q->salt & 0x3fffu) << 4) + (i++ & 15u)); };
    std::string out;
    if (!generate(g, "S", draw, &out)) return set_err(XW_ERR_INVALID_ARG, "ungrounded nonterminal");
    if (out.size() + 1 > cap) return set_err(XW_ERR_INVALID_ARG, "sentence buffer too small (%zu bytes needed)", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return (int)out.size();
}

int64_t xw_launch_count(const xw_sim* s) { return s->launches; }
int32_t xw_render_kernel(const xw_sim* s) {
    if (s->cfg.game != XW_GAME_XWORLD) return -1;
    if (s->d.vr > 0) return s->fpv_fast ? (s->fpv.regular ? 6 : 5) : 4;
    if (!s->tab.fast_ok) return 0;
    return s->render_sp ? 3 : (s->render_sb ? 1 : 2);
}

int xw_enable_timing(xw_sim* s, int32_t on) { s->timing = on != 0; return 0; }

static double avg_event_ms(xw_sim* s, std::vector<cudaEvent_t>& ev, size_t& used, int32_t reset) {
    if (!s->timing && used == 0) return -1.0;
    if (s->own_stream) cudaStreamSynchronize(s->own_stream);
    cudaDeviceSynchronize();
    double total = 0;
    size_t pairs = used / 2;
    for (size_t i = 0; i < pairs; ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]);
        total += ms;
    }
    double avg = pairs ? total / pairs : -1.0;
    if (reset) used = 0;
    return avg;
}
double xw_render_ms(xw_sim* s, int32_t reset) {
    DevGuard dev_guard(s);
    return avg_event_ms(s, s->ev, s->ev_used, reset);
}
double xw_step_reset_ms(xw_sim* s, int32_t reset) {
    DevGuard dev_guard(s);
    return avg_event_ms(s, s->ev2, s->ev2_used, reset);
}

// fix != NULL: the painter leaves `fix->reserve` SMs free, and once `fix->done` has fired the frames of the envs in the
// step's auto-reset queue are painted again (the painter in list mode) -- see step_xworld.
struct RenderFix { int reserve; cudaEvent_t done; const int32_t* list; const int32_t* count; int est; std::function<int()> after_painter;
                   bool pdl_painter = false;   // the reset kernel is launched by after_painter as the painter's programmatic dependent
                   int sweep_grid = 0;   // > 0: a second reset launch on every SM once the painter and the first one are done (what a burst left)
                   std::function<int()> between; };   // between: first-person view, called right behind the frame kernel (before the goal kernel)
static int launch_render(xw_sim* s, uint8_t* d_frames, cudaStream_t st, const RenderFix* fix = nullptr) {
    XwRender& r = s->r;
    const int K = s->cfg.context;
    const int FBo = s->C * r.OH * r.OW;  // bytes of one frame as the caller sees it
    const size_t env_stride = (size_t)K * FBo;
    if (K > 1) {  // GameSimulator::shift_context (simulator.cpp:51-60)
        if (FBo % 16) return set_err(XW_ERR_UNSUPPORTED, "context > 1 needs a frame size divisible by 16");
        k_shift_context<<<s->n_sms * 4, 256, 0, st>>>(d_frames, s->n, K, FBo, s->d.ctx_flag);
        s->launches++;
    }
    uint8_t* out = d_frames + (size_t)(K - 1) * FBo;  // newest frame last
    // --color=false: the colour frame goes to a scratch buffer, k_gray writes the caller's (xworld_simulator.cpp:529-531)
    uint8_t* dst = s->cfg.gray ? s->d_bgr : out;
    const size_t dst_stride = s->cfg.gray ? (size_t)r.FB : env_stride;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (s->timing) {
        if (s->ev_used + 2 > s->ev.size()) {
            for (int i = 0; i < 2; ++i) { cudaEvent_t ev; CUDA_TRY(cudaEventCreate(&ev)); s->ev.push_back(ev); }
        }
        e0 = s->ev[s->ev_used]; e1 = s->ev[s->ev_used + 1];
        s->ev_used += 2;
        CUDA_TRY(cudaEventRecord(e0, st));
    }
    if (s->d.vr > 0) {
        if (s->fpv_fast && dst_stride % 16 == 0) {
            if (s->fpv.regular) {
                const int nch = s->fpv_chunks, G = s->fpv.G;
                const int slots = XW_FPV_SLOTS, pbase = (s->fpv_parity ^= 1) * slots, zbase = (s->fpv_parity ^ 1) * slots;
                const bool streaming = s->fpv_streaming && nch == 1;
                if ((nch > 1 || streaming) && !s->fpv_stream) {
                    CUDA_TRY(cudaStreamCreateWithFlags(&s->fpv_stream, cudaStreamNonBlocking));
                    for (auto& e : s->fpv_ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                }
                for (int k = 0; k < nch; ++k) {
                    const int env0 = (int)((int64_t)s->n * k / nch), cnt = (int)((int64_t)s->n * (k + 1) / nch) - env0;
                    if (cnt <= 0) continue;
                    int grid = s->fpv_grid - (fix ? fix->reserve : 0);
                    const int need = (cnt + s->fpv_ng - 1) / s->fpv_ng;
                    if (grid > need) grid = need;
                    if (grid < 1) grid = 1;
                    int seq = 0;
                    if (streaming) {
                        // the goal kernel BESIDE the frame kernel: on its own stream, behind everything queued so far (the previous
                        // render's users of the list), one CTA per SM
                        seq = s->fpv_seq = s->fpv_seq % 4000 + 1;
                        CUDA_TRY(cudaEventRecord(s->fpv_ev[0], st));
                        CUDA_TRY(cudaStreamWaitEvent(s->fpv_stream, s->fpv_ev[0], 0));
                        s->fpv_stream_fn<<<s->n_sms, 512, 3 * 16384 + 128, s->fpv_stream>>>(s->fpv, dst, dst_stride, pbase, env0 * G, seq, grid * s->fpv_ng,
                                                                                              s->n * G);
                        CUDA_TRY(cudaEventRecord(s->fpv_ev[xw_sim::FPV_MAX_CHUNKS], s->fpv_stream));
                    }
                    s->fpv_cells_fn<<<grid, s->fpv_nt, s->fpv_smem, st>>>(s->d, s->fpv, dst, dst_stride, nullptr, nullptr, env0, cnt, pbase + k, env0 * G, zbase, slots, seq);
                    if (fix && fix->between) { const int rcb = fix->between(); if (rcb) return rcb; }
                    if (s->trace && s->tr[5]) CUDA_TRY(cudaEventRecord(s->tr[5], st));
                    s->launches += 2;
                    if (streaming) { CUDA_TRY(cudaStreamWaitEvent(st, s->fpv_ev[xw_sim::FPV_MAX_CHUNKS], 0)); continue; }
                    cudaStream_t gs = st;
                    if (nch > 1) {
                        gs = s->fpv_stream;
                        CUDA_TRY(cudaEventRecord(s->fpv_ev[k], st));
                        CUDA_TRY(cudaStreamWaitEvent(gs, s->fpv_ev[k], 0));
                    }
                    s->fpv_goal_fn<<<s->fpv_goal_grid, s->fpv_goal_nt, 16384 + 16, gs>>>(s->fpv, dst, dst_stride, pbase + k, env0 * G);
                }
                s->launches--;   // (the common increment below counts one)
                if (nch > 1) {
                    CUDA_TRY(cudaEventRecord(s->fpv_ev[xw_sim::FPV_MAX_CHUNKS], s->fpv_stream));
                    CUDA_TRY(cudaStreamWaitEvent(st, s->fpv_ev[xw_sim::FPV_MAX_CHUNKS], 0));
                }
            }
            else k_render_fpv<256><<<s->fpv_grid < s->n ? s->fpv_grid : s->n, 256, s->fpv_smem, st>>>(s->d, s->fpv, dst, dst_stride);
        } else {
            k_render_fpv_generic<<<s->n_sms * 8 < s->n ? s->n_sms * 8 : s->n, 256, 0, st>>>(s->d, s->fpv, dst, dst_stride);
        }
    } else if (s->tab.fast_ok) {
        const int need = (s->n + r.G - 1) / r.G;  // CTAs that get at least one env
        int grid = s->render_grid - (fix ? fix->reserve : 0);
        if (grid < 1) grid = 1;
        if (grid > need) grid = need;
        cudaLaunchConfig_t lc;
        memset(&lc, 0, sizeof lc);
        lc.gridDim = dim3(grid); lc.blockDim = dim3(r.G * r.GT); lc.dynamicSmemBytes = s->render_smem; lc.stream = st;
        cudaLaunchAttribute at[1];
        if (s->l2_window) {  // keep the tables in L2 while the frames stream through it
            at[0].id = cudaLaunchAttributeAccessPolicyWindow;
            at[0].val.accessPolicyWindow.base_ptr = s->tables;
            at[0].val.accessPolicyWindow.num_bytes = s->l2_window;
            at[0].val.accessPolicyWindow.hitRatio = s->l2_ratio;
            at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            lc.attrs = at; lc.numAttrs = 1;
        }
        CUDA_TRY(cudaLaunchKernelEx(&lc, s->render_fn, s->d, r, dst, dst_stride));
    } else {
        k_render_generic<<<s->n_sms * 8, 256, 0, st>>>(s->d, r, dst, dst_stride);
    }
    s->launches++;
    // kernel timing: the event behind the frame kernel(s) -- not the re-paint of the reset queue, not the gray pass.  When the reset
    // kernel is the painter's programmatic dependent nothing may sit between the two launches: the event then follows the reset
    // launch and measures painter || reset (the reset ends first on the bench workloads, so it is the painter's time; never less)
    const bool e1_late = s->timing && fix && fix->pdl_painter;
    if (s->timing && !e1_late) CUDA_TRY(cudaEventRecord(e1, st));
    if (fix) {
        if (fix->after_painter) {
            const int rc2 = fix->after_painter();  // (the painter is queued: now the reset launch -- its dependent, or on its own stream, then `done`)
            if (rc2) return rc2;
        }
        if (e1_late) CUDA_TRY(cudaEventRecord(e1, st));
        if (fix->done) CUDA_TRY(cudaStreamWaitEvent(st, fix->done, 0));
        if (fix->sweep_grid > 0) {
            k_reset_list<<<fix->sweep_grid, 128, 0, st>>>(s->d, fix->list, fix->count, 0);
            s->launches++;
            if (s->d.vr > 0) {
                k_fpv_warp_goals<<<s->n_sms * 8 < s->n ? s->n_sms * 8 : s->n, 256, 0, st>>>(s->d, s->fpv, nullptr, fix->list, fix->count);
                s->launches++;
            }
        }
        if (s->d.vr > 0) {  // first-person view: the frame kernel and the goal kernel again, over the queue (its own goal-list slot)
            const int slot = s->fpv_parity * XW_FPV_SLOTS + xw_sim::FPV_MAX_CHUNKS;   // (the main pass zeroed nothing of this half)
            int blocks = (fix->est + s->fpv_ng - 1) / s->fpv_ng;
            if (blocks > s->fpv_grid) blocks = s->fpv_grid;
            s->fpv_cells_fn<<<blocks, s->fpv_nt, s->fpv_smem, st>>>(s->d, s->fpv, dst, dst_stride, fix->list, fix->count, 0, 0, slot, 0, 0, 0, 0);
            blocks = 2 * fix->est < s->fpv_goal_grid ? 2 * fix->est : s->fpv_goal_grid;
            s->fpv_goal_fn<<<blocks, s->fpv_goal_nt, 16384 + 16, st>>>(s->fpv, dst, dst_stride, slot, 0);
            s->launches += 2;
        } else {
            // the painter again, in list mode: about one env per warp group at the expected queue length
            int blocks = (fix->est + r.G - 1) / r.G;
            if (blocks > s->render_grid) blocks = s->render_grid;
            XwRender rl = r;
            rl.env_list = fix->list; rl.env_count = fix->count;
            s->render_list_fn<<<blocks, r.G * r.GT, s->render_smem, st>>>(s->d, rl, dst, dst_stride);
            s->launches++;
        }
    }
    if (s->cfg.gray) {
        k_gray<<<s->n_sms * 8, 256, 0, st>>>(s->d_bgr, out, s->n, r.OH * r.OW, (size_t)r.FB, env_stride);
        s->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int xw_render(xw_sim* s, uint8_t* d_frames, void* stream) {
    DevGuard dev_guard(s);
    if (s->cfg.game != XW_GAME_XWORLD) return set_err(XW_ERR_UNSUPPORTED, "xw_render: xworld only");
    if (!d_frames) return set_err(XW_ERR_INVALID_ARG, "null frames");
    return launch_render(s, d_frames, pick_stream(s, stream));
}

int xw_reset(xw_sim* s, const uint8_t* d_mask, void* stream) {
    DevGuard dev_guard(s);
    cudaStream_t st = pick_stream(s, stream);
    if (s->cfg.game == XW_GAME_XWORLD) {
        k_reset<<<(s->n + 127) / 128, 128, 0, st>>>(s->d, d_mask);
        s->launches++;
        if (s->d.vr > 0) {  // the new episodes' goal icons, warped once (xitem.cpp:47-60)
            k_fpv_warp_goals<<<s->n_sms * 8 < s->n ? s->n_sms * 8 : s->n, 256, 0, st>>>(s->d, s->fpv, d_mask, nullptr, nullptr);
            s->launches++;
        }
    } else if (s->cfg.game == XW_GAME_SIMPLE_RACE) {
        k_race_reset<<<(s->n + 255) / 256, 256, 0, st>>>(s->race, d_mask);
        s->launches++;
    } else {
        return set_err(XW_ERR_UNSUPPORTED, "simple_game runs on the host: use xw_reset_host");
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// One SimulatorInterface::take_actions of the xworld batch on stream st: k_step, the auto-reset launch, the frames.
// ev_step (may be NULL) is recorded when reward / game_over / the invalid-action count are final, i.e. after k_step.
//
// Reset beside the painter.  The auto-reset launch is latency, not work: a few hundred envs per step, one warp each, 80-190 us
// of one long dependent instruction stream per env, during which 99 % of the GPU idles and the painter waits.  The frames of the
// envs that do NOT reset depend on k_step only, so (fully observed view, context 1, shared-memory painter):
//   st:            k_step -> painter on (SMs - reserve) CTAs, all envs -> [wait] -> the painter in list mode: the queue's envs again
//   reset_stream:  [k_step done] -> k_reset_list on the SMs the painter left free -> [done]
// The painter reads the queued envs' old (possibly half-rewritten: always valid cell codes) state; those frames are overwritten
// by the second, tiny launch.  `reserve` follows the queue lengths of the last steps (a pinned counter, read without waiting).
static int step_xworld(xw_sim* s, const int32_t* d_actions, int act_rep, float* d_reward, int32_t* d_game_over, uint8_t* d_frames,
                       cudaStream_t st, cudaEvent_t ev_step) {
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (s->timing) {
        if (s->ev2_used + 2 > s->ev2.size())
            for (int i = 0; i < 2; ++i) { cudaEvent_t ev; CUDA_TRY(cudaEventCreate(&ev)); s->ev2.push_back(ev); }
        t0 = s->ev2[s->ev2_used]; t1 = s->ev2[s->ev2_used + 1];
        s->ev2_used += 2;
        CUDA_TRY(cudaEventRecord(t0, st));
    }
    if (s->trace) {
        if (!s->tr[0]) for (auto& e : s->tr) CUDA_TRY(cudaEventCreate(&e));
        if (s->trace_steps > 0 && s->trace_steps % 16 == 8 && s->reset_stream) {  // report an earlier step (complete by now)
            cudaEventSynchronize(s->tr[4]); cudaEventSynchronize(s->tr[3]);
            float a = 0, b = 0, c2 = 0, d2 = 0, f2 = 0;
            if (s->d.vr > 0 && cudaEventElapsedTime(&f2, s->tr[0], s->tr[5]) == cudaSuccess) fprintf(stderr, "XW_TRACE fpv frame kernel end %.1f us (a later step's)\n", f2 * 1e3);
            cudaEventElapsedTime(&a, s->tr[0], s->tr[1]); cudaEventElapsedTime(&b, s->tr[0], s->tr[2]);
            cudaEventElapsedTime(&c2, s->tr[0], s->tr[3]); cudaEventElapsedTime(&d2, s->tr[0], s->tr[4]);
            fprintf(stderr, "XW_TRACE step %d: k_step end %.1f us | painter end %.1f | reset end %.1f | re-paint end %.1f (reset_avg %.0f)\n",
                    s->trace_steps, a * 1e3, b * 1e3, c2 * 1e3, d2 * 1e3, s->reset_avg);
        }
    }
    const bool tracing = s->trace && (s->trace_steps++ % 16 == 7);
    if (tracing) CUDA_TRY(cudaEventRecord(s->tr[0], st));
    k_step<<<(s->n + 255) / 256, 256, 0, st>>>(s->d, d_actions, act_rep, d_reward, d_game_over, s->step_parity);
    s->launches++;
    if (tracing) CUDA_TRY(cudaEventRecord(s->tr[1], st));
    if (ev_step) CUDA_TRY(cudaEventRecord(ev_step, st));
    const int32_t* q_list = s->d.reset_list;
    const int32_t* q_count = s->d.reset_count + s->step_parity;
    const int parity = s->step_parity;
    s->step_parity ^= 1;
    const bool fpv_ok = s->d.vr > 0 && s->fpv_fast && s->fpv.regular && !s->cfg.gray && ((size_t)s->cfg.context * s->C * s->r.OH * s->r.OW) % 16 == 0;
    const bool overlap = s->overlap_reset && s->cfg.auto_reset && d_frames && s->cfg.context == 1 &&
                         (s->d.vr == 0 ? (s->tab.fast_ok && s->render_list_fn != nullptr) : fpv_ok);
    if (overlap) {
        if (!s->reset_stream) {
            // (highest priority: when CTA slots come free -- the first-person view's frame kernel ending -- the reset chain's CTAs
            //  are placed before the goal kernel's, which is queued on the render stream)
            int prio_lo = 0, prio_hi = 0;
            CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            CUDA_TRY(cudaStreamCreateWithPriority(&s->reset_stream, cudaStreamNonBlocking, prio_hi));
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_a, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_b, cudaEventDisableTiming));
            CUDA_TRY(cudaMallocHost((void**)&s->h_reset_cnt, 2 * sizeof(int32_t)));
            s->h_reset_cnt[0] = s->h_reset_cnt[1] = -1;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->reset_ctas_per_sm, k_reset_list, 128, 0));
            if (s->reset_ctas_per_sm < 1) s->reset_ctas_per_sm = 1;
        }
        const int32_t seen = s->h_reset_cnt[parity ^ 1];  // the previous step's queue length, if its copy has landed
        if (seen >= 0) s->reset_avg = 0.75f * s->reset_avg + 0.25f * (float)seen;
        // one warp per queued env, reset_ctas_per_sm CTAs of 4 warps per reserved SM, about three rounds inside a render
        static const int rounds = [] { const char* e = getenv("XW_RESET_ROUNDS"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 16 ? v : 3; }();
        int reserve = (int)(s->reset_avg / (float)(s->reset_ctas_per_sm * 4 * rounds)) + 1;
        if (reserve > s->n_sms / 8) reserve = s->n_sms / 8;
        const bool fpv = s->d.vr > 0;
        // the CTAs that land on the SMs the painter leaves free drain a normal queue; the others -- a CTA for every slot of the
        // device -- start when the painter's CTAs leave and find the queue empty, or finish a burst (k_reset_list).  (Measured: the
        // empty CTAs cost a step nothing, 0.3315 ms with 444 and with 172 of them.)
        int reset_grid = s->n_sms * s->reset_ctas_per_sm;
        if (reset_grid > (s->n + 3) / 4) reset_grid = (s->n + 3) / 4;
        static const int fdbg = [] { const char* e = getenv("XW_FPV_DBG"); return e ? atoi(e) : 0; }();
        if (fpv && (fdbg & 4)) reserve = 0;
        if (t1) CUDA_TRY(cudaEventRecord(t1, st));
        RenderFix fix = {reserve, s->ev_b, q_list, q_count, (int)(s->reset_avg * 1.5f) + 16, nullptr};
        if (!fpv && s->reset_pdl && !s->trace) {
            // The reset kernel as the painter's PROGRAMMATIC DEPENDENT in the same stream: it may start as soon as every painter
            // CTA has executed griddepcontrol.launch_dependents (the first instruction of k_render_sp), i.e. when the painter holds
            // its SMs, and it never waits for the painter.  With the reset on a second stream behind an event both kernels become
            // runnable together when k_step ends; when the reset's CTAs were placed first -- one per SM, the hardware spreads them --
            // that many painter CTAs (a whole SM each) had to wait for a reset round before they could start: 100 us per step in the
            // synchronous host-buffer call, where the GPU is idle when the step begins (profiles/r02_summary.md).
            fix.done = nullptr;
            fix.pdl_painter = true;
            fix.after_painter = [s, reset_grid, q_list, q_count, parity, tracing, st]() -> int {
                cudaLaunchConfig_t lc;   // (nothing may sit between the painter and its dependent in the stream)
                memset(&lc, 0, sizeof lc);
                lc.gridDim = dim3(reset_grid); lc.blockDim = dim3(128); lc.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                lc.attrs = at; lc.numAttrs = 1;
                CUDA_TRY(cudaLaunchKernelEx(&lc, k_reset_list, s->d, q_list, q_count, 0));
                s->launches++;
                if (tracing) { CUDA_TRY(cudaEventRecord(s->tr[2], st)); CUDA_TRY(cudaEventRecord(s->tr[3], st)); }   // (= painter and reset both done)
                CUDA_TRY(cudaMemcpyAsync(s->h_reset_cnt + parity, q_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
                return 0;
            };
            const int rc3 = launch_render(s, d_frames, st, &fix);
            if (tracing) CUDA_TRY(cudaEventRecord(s->tr[4], st));
            return rc3;
        }
        if (fpv && s->reset_pdl && s->fpv_chunks == 1) {
            // first-person view: the reset kernel is the FRAME kernel's programmatic dependent (same reasons); the new episodes' goal
            // icons are warped on the reset stream once frame kernel and reset are done, beside the main pass's goal kernel; the
            // re-paint of the queue waits for both
            fix.between = [s, reset_grid, q_list, q_count, parity, st]() -> int {
                cudaLaunchConfig_t lc;
                memset(&lc, 0, sizeof lc);
                lc.gridDim = dim3(reset_grid); lc.blockDim = dim3(128); lc.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                lc.attrs = at; lc.numAttrs = 1;
                CUDA_TRY(cudaLaunchKernelEx(&lc, k_reset_list, s->d, q_list, q_count, 0));
                CUDA_TRY(cudaEventRecord(s->ev_a, st));
                CUDA_TRY(cudaStreamWaitEvent(s->reset_stream, s->ev_a, 0));
                const int g2 = (int)(s->reset_avg * 1.25f) * s->fpv.G + 8;
                k_fpv_warp_goals<<<g2 < s->n_sms * 8 ? g2 : s->n_sms * 8, 256, 0, s->reset_stream>>>(s->d, s->fpv, nullptr, q_list, q_count);
                CUDA_TRY(cudaEventRecord(s->ev_b, s->reset_stream));
                CUDA_TRY(cudaMemcpyAsync(s->h_reset_cnt + parity, q_count, sizeof(int32_t), cudaMemcpyDeviceToHost, s->reset_stream));
                s->launches += 2;
                return 0;
            };
            const int rc3 = launch_render(s, d_frames, st, &fix);
            if (tracing) { CUDA_TRY(cudaEventRecord(s->tr[2], st)); CUDA_TRY(cudaEventRecord(s->tr[3], st)); CUDA_TRY(cudaEventRecord(s->tr[4], st)); }
            return rc3;
        }
        CUDA_TRY(cudaEventRecord(s->ev_a, st));
        CUDA_TRY(cudaStreamWaitEvent(s->reset_stream, s->ev_a, 0));
        // (the painter first: its CTAs take their SMs before the reset launch, which waits for an event, can be placed)
        const int small_grid = reserve * s->reset_ctas_per_sm;   // two-stream form: the launch beside the painter stays on the reserved SMs
        fix.sweep_grid = reset_grid;
        fix.after_painter = [s, small_grid, fpv, q_list, q_count, parity, tracing, st]() -> int {
            if (tracing) CUDA_TRY(cudaEventRecord(s->tr[2], st));
            static const int fdbg = [] { const char* e = getenv("XW_FPV_DBG"); return e ? atoi(e) : 0; }();
            if (!(fpv && (fdbg & 2))) k_reset_list<<<small_grid, 128, 0, s->reset_stream>>>(s->d, q_list, q_count, 8);
            s->launches++;
            if (fpv && !(fdbg & 1)) {  // the new episodes' goal icons (one CTA per queued env and goal)
                const int g2 = (int)(s->reset_avg * 1.25f) * s->fpv.G + 8;
                k_fpv_warp_goals<<<g2 < s->n_sms * 8 ? g2 : s->n_sms * 8, 256, 0, s->reset_stream>>>(s->d, s->fpv, nullptr, q_list, q_count);
                s->launches++;
            }
            if (tracing) CUDA_TRY(cudaEventRecord(s->tr[3], s->reset_stream));
            CUDA_TRY(cudaEventRecord(s->ev_b, s->reset_stream));
            static const bool nocnt = [] { const char* e = getenv("XW_OVERLAP_NOCNT"); return e && atoi(e) != 0; }();
            if (!nocnt) CUDA_TRY(cudaMemcpyAsync(s->h_reset_cnt + parity, q_count, sizeof(int32_t), cudaMemcpyDeviceToHost, s->reset_stream));
            CUDA_TRY(cudaGetLastError());
            return 0;
        };
        const int rc3 = launch_render(s, d_frames, st, &fix);
        if (tracing) CUDA_TRY(cudaEventRecord(s->tr[4], st));
        return rc3;
    }
    if (s->cfg.auto_reset) {
        const int want = (s->n + 3) / 4, cap = s->n_sms * 16;  // CTAs of 4 warps: one warp per queued env, grid-stride past the cap
        k_reset_list<<<want < cap ? want : cap, 128, 0, st>>>(s->d, q_list, q_count, 0);
        s->launches++;
        if (s->d.vr > 0) {
            k_fpv_warp_goals<<<s->n_sms * 8 < s->n ? s->n_sms * 8 : s->n, 256, 0, st>>>(s->d, s->fpv, nullptr, q_list, q_count);
            s->launches++;
        }
    }
    if (t1) CUDA_TRY(cudaEventRecord(t1, st));
    CUDA_TRY(cudaGetLastError());
    if (d_frames) return launch_render(s, d_frames, st);
    return 0;
}

int xw_step(xw_sim* s, const int32_t* d_actions, int32_t act_rep, float* d_reward, int32_t* d_game_over,
            uint8_t* d_frames, void* stream) {
    DevGuard dev_guard(s);
    if (!d_actions || !d_reward || !d_game_over) return set_err(XW_ERR_INVALID_ARG, "null buffer");
    if (act_rep < 1) return set_err(XW_ERR_INVALID_ARG, "act_rep must be >= 1");
    cudaStream_t st = pick_stream(s, stream);
    if (s->cfg.game == XW_GAME_XWORLD) return step_xworld(s, d_actions, act_rep, d_reward, d_game_over, d_frames, st, nullptr);
    if (s->cfg.game == XW_GAME_SIMPLE_RACE) {
        // (GameSimulator::take_actions repeats the action act_rep times: inside the kernel, one launch per call)
        k_race_step<<<(s->n + 255) / 256, 256, 0, st>>>(s->race, d_actions, xw_num_actions(s), act_rep, d_reward, d_game_over, s->race_error);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        if (d_frames)  // "screen" of simple_race = the 4-float state vector
            CUDA_TRY(cudaMemcpyAsync(d_frames, s->race.state, sizeof(float) * 4 * s->n, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    return set_err(XW_ERR_UNSUPPORTED, "simple_game runs on the host: use xw_step_host");
}

// K consecutive take_actions calls per env, no frames: d_actions / d_reward / d_game_over are [K][n_envs] (step-major).
// simple_race: one launch (k_race_step_seq); xworld: K step (+ auto-reset) launches queued back to back.  Results are
// those of K xw_step calls with d_frames == NULL.
int xw_step_seq(xw_sim* s, const int32_t* d_actions, int32_t k_steps, int32_t act_rep, float* d_reward, int32_t* d_game_over, void* stream) {
    DevGuard dev_guard(s);
    if (!d_actions || !d_reward || !d_game_over) return set_err(XW_ERR_INVALID_ARG, "null buffer");
    if (act_rep < 1 || k_steps < 1) return set_err(XW_ERR_INVALID_ARG, "act_rep and k_steps must be >= 1");
    cudaStream_t st = pick_stream(s, stream);
    if (s->cfg.game == XW_GAME_SIMPLE_RACE) {
        k_race_step_seq<<<(s->n + 255) / 256, 256, 0, st>>>(s->race, d_actions, xw_num_actions(s), act_rep, k_steps, d_reward, d_game_over, s->race_error);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    if (s->cfg.game == XW_GAME_XWORLD) {
        for (int k = 0; k < k_steps; ++k) {
            const int rc = step_xworld(s, d_actions + (size_t)k * s->n, act_rep, d_reward + (size_t)k * s->n, d_game_over + (size_t)k * s->n, nullptr, st, nullptr);
            if (rc) return rc;
        }
        return 0;
    }
    return set_err(XW_ERR_UNSUPPORTED, "simple_game runs on the host: use xw_step_host");
}

static int invalid_status(xw_sim* s);
static int ensure_staging(xw_sim* s, bool frames) {
    if (!s->h_act) {
        CUDA_TRY(cudaMallocHost((void**)&s->h_act, sizeof(int32_t) * s->n));
        CUDA_TRY(cudaMallocHost((void**)&s->h_over, sizeof(int32_t) * s->n));
        CUDA_TRY(cudaMallocHost((void**)&s->h_rew, sizeof(float) * s->n));
        CUDA_TRY(cudaMallocHost((void**)&s->h_invalid, sizeof(int32_t)));
        *s->h_invalid = 0;
        int rc = dalloc(s, &s->d_act, (size_t)s->n) | dalloc(s, &s->d_over, (size_t)s->n) | dalloc(s, &s->d_rew, (size_t)s->n) |
                 dalloc(s, &s->d_mask, (size_t)s->n);
        if (rc) return rc;
        if (s->cfg.game == XW_GAME_XWORLD) { s->d.stage_rew = s->d_rew; s->d.stage_over = s->d_over; }
    }
    if (frames && !s->d_frames) return dalloc(s, &s->d_frames, (size_t)s->n * xw_frame_bytes(s));
    return 0;
}

// ---- simple_game on the host (SimpleGameEngine, simple_game_simulator.cpp:24-76) ----
static void sg_reset(const xw_config& c, SimpleGameEnv& g) {
    g.cur_pos = c.array_size / 2;
    g.num_steps = 0;
    g.state.assign(c.array_size, 0);
    g.rewards.assign(c.array_size, 0.f);
    g.state[g.cur_pos] = 1;
    g.rewards[c.array_size - 1] = 4.0f / 2;  // DEST_REWARD / 2
    g.rewards[0] = 4.0f;
}
static bool sg_over(const SimpleGameEnv& g) { return g.cur_pos <= 0 || g.cur_pos >= (int)g.state.size() - 1; }
static float sg_reward(SimpleGameEnv& g) {
    float reward = -0.1f;  // MOVE_REWARD
    if (g.cur_pos >= 0 && g.cur_pos < (int)g.state.size() && g.rewards[g.cur_pos] != 0.0) {
        reward = g.rewards[g.cur_pos];
        g.rewards[g.cur_pos] = 0.0;
    }
    return reward;
}
static float sg_act(SimpleGameEnv& g, int a) {
    if (sg_over(g)) return sg_reward(g);
    g.state[g.cur_pos] = 0;
    if (a == 0) --g.cur_pos; else ++g.cur_pos;
    if (g.cur_pos >= 0 && g.cur_pos < (int)g.state.size()) g.state[g.cur_pos] = 1;
    return sg_reward(g);
}

int xw_reset_host(xw_sim* s, const uint8_t* h_mask, uint8_t* h_frames) {
    DevGuard dev_guard(s);
    if (s->cfg.game == XW_GAME_SIMPLE_GAME) {
        for (int i = 0; i < s->n; ++i) {
            if (h_mask && !h_mask[i]) continue;
            sg_reset(s->cfg, s->sg[i]);
        }
        if (h_frames)
            for (int i = 0; i < s->n; ++i) memcpy(h_frames + (size_t)i * s->cfg.array_size, s->sg[i].state.data(), s->cfg.array_size);
        return 0;
    }
    int rc = ensure_staging(s, h_frames != nullptr);
    if (rc) return rc;
    cudaStream_t st = s->own_stream;
    if (h_mask) CUDA_TRY(cudaMemcpyAsync(s->d_mask, h_mask, s->n, cudaMemcpyHostToDevice, st));
    rc = xw_reset(s, h_mask ? s->d_mask : nullptr, st);
    if (rc) return rc;
    if (h_frames) {
        if (s->cfg.game == XW_GAME_XWORLD) {
            if (s->cfg.context > 1 && !h_mask)  // init_screen: zero-filled context, newest frame last
                CUDA_TRY(cudaMemsetAsync(s->d_frames, 0, (size_t)s->n * xw_frame_bytes(s), st));
            rc = launch_render(s, s->d_frames, st);
            if (rc) return rc;
        } else {
            CUDA_TRY(cudaMemsetAsync(s->d_frames, 0, (size_t)s->n * xw_frame_bytes(s), st));
        }
        CUDA_TRY(cudaMemcpyAsync(h_frames, s->d_frames, (size_t)s->n * xw_frame_bytes(s), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int xw_step_host(xw_sim* s, const int32_t* h_actions, int32_t act_rep, float* h_reward, int32_t* h_game_over, uint8_t* h_frames) {
    DevGuard dev_guard(s);
    if (!h_actions || !h_reward || !h_game_over) return set_err(XW_ERR_INVALID_ARG, "null buffer");
    if (s->cfg.game == XW_GAME_SIMPLE_GAME) {
        for (int i = 0; i < s->n; ++i) {
            SimpleGameEnv& g = s->sg[i];
            if (h_actions[i] == XW_ACTION_NONE) continue;
            if (h_actions[i] < 0 || h_actions[i] >= 2) return set_err(XW_ERR_INVALID_ACTION, "undefined action_id: %d", h_actions[i]);
            float r = 0;
            g.num_steps++;
            for (int rep = 0; rep < act_rep; ++rep) r += sg_act(g, h_actions[i]);
            h_reward[i] = r;
            int over = 0;
            if (s->cfg.max_steps > 0 && g.num_steps >= s->cfg.max_steps) over |= XW_MAX_STEP;
            if (sg_over(g)) over |= XW_SUCCESS;  // SimpleGame::game_over, simple_game_simulator.cpp:83-85
            h_game_over[i] = over;
            if (h_frames) memcpy(h_frames + (size_t)i * s->cfg.array_size, g.state.data(), s->cfg.array_size);
        }
        return 0;
    }
    int rc = ensure_staging(s, h_frames != nullptr);
    if (rc) return rc;
    cudaStream_t st = s->own_stream;
    memcpy(s->h_act, h_actions, sizeof(int32_t) * s->n);
    CUDA_TRY(cudaMemcpyAsync(s->d_act, s->h_act, sizeof(int32_t) * s->n, cudaMemcpyHostToDevice, st));
    rc = xw_step(s, s->d_act, act_rep, s->d_rew, s->d_over, h_frames ? s->d_frames : nullptr, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(s->h_rew, s->d_rew, sizeof(float) * s->n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(s->h_over, s->d_over, sizeof(int32_t) * s->n, cudaMemcpyDeviceToHost, st));
    if (s->cfg.game == XW_GAME_XWORLD) CUDA_TRY(cudaMemcpyAsync(s->h_invalid, s->d.n_invalid, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (h_frames) CUDA_TRY(cudaMemcpyAsync(h_frames, s->d_frames, (size_t)s->n * xw_frame_bytes(s), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    memcpy(h_reward, s->h_rew, sizeof(float) * s->n);   // (slots of XW_ACTION_NONE envs: their last values; 0 / alive after a reset)
    memcpy(h_game_over, s->h_over, sizeof(int32_t) * s->n);
    return invalid_status(s);
}

// SURVEY §8b "per-env flag + status": the calls that hand results to the host also report whether any env of this
// step was given an action outside [0, num_actions) (the reference CHECK-aborts, xworld_simulator.cpp:254)
static int invalid_status(xw_sim* s) {
    if (s->cfg.game != XW_GAME_XWORLD || !s->h_invalid) return 0;
    const int32_t total = *s->h_invalid, fresh = total - s->invalid_seen;
    s->invalid_seen = total;
    return fresh > 0 ? set_err(XW_ERR_INVALID_ACTION, "invalid action for %d env(s): left untouched and flagged (xw_error_flags)", fresh) : 0;
}

// Is the caller's buffer page-locked (then it is used in place)?  Asked on every call (about a microsecond each): a cache of
// addresses would mis-classify a buffer that was freed and whose address was handed out again as pageable memory.
static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// Host actions in, host reward / game_over out, frames stay on the device.  Page-locked caller buffers are used
// in place (no staging copy); the reward / game_over read-back runs on a second stream as soon as the step and
// reset kernels are done, i.e. under the render kernel, so the call returns when the frames are complete.
static int step_hd(xw_sim* s, const int32_t* h_actions, int32_t act_rep, float* h_reward, int32_t* h_game_over, uint8_t* d_frames,
                   bool wait_frames);
int xw_step_hd(xw_sim* s, const int32_t* h_actions, int32_t act_rep, float* h_reward, int32_t* h_game_over, uint8_t* d_frames) {
    return step_hd(s, h_actions, act_rep, h_reward, h_game_over, d_frames, true);
}
// The pipelined form: returns when reward / game_over are on the host; the render kernel of this step may still be
// running on the handle's stream.  Whatever reads the frames must be ordered after it: xw_wait_frames(sim, stream)
// makes a consumer stream wait, xw_sync(sim) the host.  The next call's step kernel is stream-ordered behind the
// render, so the GPU goes from one step into the next without waiting for the host.
int xw_step_hd_async(xw_sim* s, const int32_t* h_actions, int32_t act_rep, float* h_reward, int32_t* h_game_over, uint8_t* d_frames) {
    return step_hd(s, h_actions, act_rep, h_reward, h_game_over, d_frames, false);
}
int xw_wait_frames(xw_sim* s, void* stream) {
    DevGuard dev_guard(s);
    if (!s->ev_frames) CUDA_TRY(cudaEventCreateWithFlags(&s->ev_frames, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(s->ev_frames, s->own_stream));
    CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, s->ev_frames, 0));
    return 0;
}
int xw_sync(xw_sim* s) {
    DevGuard dev_guard(s);
    if (s->own_stream) CUDA_TRY(cudaStreamSynchronize(s->own_stream));
    if (s->copy_stream) CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
    if (s->reset_stream) CUDA_TRY(cudaStreamSynchronize(s->reset_stream));
    return 0;
}
static int step_hd(xw_sim* s, const int32_t* h_actions, int32_t act_rep, float* h_reward, int32_t* h_game_over, uint8_t* d_frames,
                   bool wait_frames) {
    DevGuard dev_guard(s);
    if (!h_actions || !h_reward || !h_game_over) return set_err(XW_ERR_INVALID_ARG, "null buffer");
    if (s->cfg.game == XW_GAME_SIMPLE_GAME) return set_err(XW_ERR_UNSUPPORTED, "simple_game has no device frames");
    int rc = ensure_staging(s, false);
    if (rc) return rc;
    cudaStream_t st = s->own_stream;
    struct timespec ts0, ts1, ts2;
    if (s->trace) {
        clock_gettime(CLOCK_MONOTONIC, &ts0);
        if (!s->tr_h0) { CUDA_TRY(cudaEventCreate(&s->tr_h0)); CUDA_TRY(cudaEventCreate(&s->tr_h1)); }
        if (wait_frames) CUDA_TRY(cudaEventRecord(s->tr_h0, st));
    }
    const bool pin_a = is_pinned(h_actions), pin_r = is_pinned(h_reward), pin_o = is_pinned(h_game_over);
    const int32_t* src_a = h_actions;
    if (!pin_a) { memcpy(s->h_act, h_actions, sizeof(int32_t) * s->n); src_a = s->h_act; }
    // XW_E2E_ZEROCOPY=1: the step kernel reads a page-locked action buffer over PCIe itself, no copy operation in front of it.
    // Round 1 measured +0.9 % with it; with the reset launch running beside the painter it costs 20 % of the synchronous
    // end-to-end rate (profiles/r02_summary.md, r02j: 141 M vs 177 M env-steps/s), so the H2D copy is the default again.
    static const bool zero_copy = [] { const char* e = getenv("XW_E2E_ZEROCOPY"); return e && atoi(e) != 0; }();
    const int32_t* dev_a = s->d_act;
    void* mapped = nullptr;
    const bool split = s->cfg.game == XW_GAME_XWORLD && d_frames != nullptr;
    if (split && !s->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&s->ev_step, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&s->ev_copy, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&s->ev_h2d, cudaEventDisableTiming));
    }
    if (split && !wait_frames) {
        // pipelined: the previous step's render kernel is probably still running on `st`; the actions go up on the copy
        // stream under it, and the step kernel (queued behind the render) finds them in HBM
        CUDA_TRY(cudaMemcpyAsync(s->d_act, src_a, sizeof(int32_t) * s->n, cudaMemcpyHostToDevice, s->copy_stream));
        CUDA_TRY(cudaEventRecord(s->ev_h2d, s->copy_stream));
        CUDA_TRY(cudaStreamWaitEvent(st, s->ev_h2d, 0));
    } else if (zero_copy && pin_a && cudaHostGetDevicePointer(&mapped, (void*)src_a, 0) == cudaSuccess && mapped) {
        dev_a = (const int32_t*)mapped;
    } else {
        cudaGetLastError();
        CUDA_TRY(cudaMemcpyAsync(s->d_act, src_a, sizeof(int32_t) * s->n, cudaMemcpyHostToDevice, st));
    }
    cudaStream_t cs = st;
    if (split) {  // reward / game_over go back on the copy stream as soon as k_step is done, under the reset and render kernels
        if (act_rep < 1) return set_err(XW_ERR_INVALID_ARG, "act_rep must be >= 1");
        cs = s->copy_stream;
        rc = step_xworld(s, dev_a, act_rep, s->d_rew, s->d_over, d_frames, st, s->ev_step);
        if (rc) return rc;
        CUDA_TRY(cudaStreamWaitEvent(cs, s->ev_step, 0));
    } else {
        rc = xw_step(s, dev_a, act_rep, s->d_rew, s->d_over, d_frames, st);
        if (rc) return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(pin_r ? h_reward : s->h_rew, s->d_rew, sizeof(float) * s->n, cudaMemcpyDeviceToHost, cs));
    CUDA_TRY(cudaMemcpyAsync(pin_o ? h_game_over : s->h_over, s->d_over, sizeof(int32_t) * s->n, cudaMemcpyDeviceToHost, cs));
    if (s->cfg.game == XW_GAME_XWORLD) CUDA_TRY(cudaMemcpyAsync(s->h_invalid, s->d.n_invalid, sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
    if (split) {
        if (wait_frames) {
            CUDA_TRY(cudaEventRecord(s->ev_copy, cs));  // one host wait for both streams
            CUDA_TRY(cudaStreamWaitEvent(st, s->ev_copy, 0));
        }
    }
    if (s->trace) { clock_gettime(CLOCK_MONOTONIC, &ts1); if (wait_frames) CUDA_TRY(cudaEventRecord(s->tr_h1, st)); }
    CUDA_TRY(cudaStreamSynchronize(split && !wait_frames ? cs : st));
    if (s->trace) {
        clock_gettime(CLOCK_MONOTONIC, &ts2);
        if (wait_frames) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, s->tr_h0, s->tr_h1);
            if (s->tr[0] && cudaEventElapsedTime(&b, s->tr_h0, s->tr[0]) == cudaSuccess) s->h2d_us += b * 1e3;
            s->gpu_span_us += a * 1e3;
        }
        s->host_enq_us += (ts1.tv_sec - ts0.tv_sec) * 1e6 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-3;
        s->host_sync_us += (ts2.tv_sec - ts1.tv_sec) * 1e6 + (ts2.tv_nsec - ts1.tv_nsec) * 1e-3;
        if (++s->host_n % 32 == 0) {
            fprintf(stderr, "XW_TRACE step_hd: host enqueue %.1f us, wait %.1f us per call (last 32); GPU first-to-last op %.1f us (last k_step-start sample at +%.1f)\n",
                    s->host_enq_us / 32, s->host_sync_us / 32, s->gpu_span_us / 32, s->h2d_us / 32);
            s->host_enq_us = s->host_sync_us = s->gpu_span_us = s->h2d_us = 0;
        }
    }
    if (!pin_r) memcpy(h_reward, s->h_rew, sizeof(float) * s->n);
    if (!pin_o) memcpy(h_game_over, s->h_over, sizeof(int32_t) * s->n);
    return invalid_status(s);
}

int xw_num_steps(xw_sim* s, int64_t* h) {
    DevGuard dev_guard(s);
    if (s->cfg.game == XW_GAME_SIMPLE_GAME) { for (int i = 0; i < s->n; ++i) h[i] = s->sg[i].num_steps; return 0; }
    std::vector<int32_t> tmp(s->n);
    const int32_t* src = s->cfg.game == XW_GAME_XWORLD ? s->d.num_steps : s->race.steps;
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(tmp.data(), src, sizeof(int32_t) * s->n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < s->n; ++i) h[i] = tmp[i];
    return 0;
}

// ---- named state fields ----
struct FieldRef { void* ptr; size_t elem; size_t per_env; bool goal_major; };
static bool find_field(xw_sim* s, const char* name, FieldRef* f) {
    std::string k(name);
    if (s->cfg.game == XW_GAME_XWORLD) {
        XwDev& d = s->d;
        struct { const char* n; void* p; size_t elem, per; bool gm; } tbl[] = {
            {"grid", d.grid, 1, (size_t)d.CS, false},
            {"agent_x", d.agent_x, 1, 1, false}, {"agent_y", d.agent_y, 1, 1, false}, {"facing", d.facing, 1, 1, false},
            {"task", d.task, 1, 1, false}, {"stage", d.stage, 1, 1, false}, {"event", d.event, 1, 1, false},
            {"action_success", d.succ, 1, 1, false}, {"target_mask", d.tmask, 1, 1, false},
            {"aux0", d.aux0, 1, 1, false}, {"aux1", d.aux1, 1, 1, false}, {"aux2", d.aux2, 1, 1, false},
            {"goal_x", d.goal_x, 1, XW_MAX_GOALS, true}, {"goal_y", d.goal_y, 1, XW_MAX_GOALS, true},
            {"goal_icon", d.goal_icon, 4, XW_MAX_GOALS, true}, {"goal_name", d.goal_name, 4, XW_MAX_GOALS, true},
            {"goal_yaw", d.goal_yaw, 2, XW_MAX_GOALS, true}, {"goal_scale", d.goal_scale, 8, XW_MAX_GOALS, true},
            {"goal_offset", d.goal_offset, 8, XW_MAX_GOALS, true},
            {"steps_in_task", d.steps_in_task, 4, 1, false}, {"num_steps", d.num_steps, 4, 1, false},
            {"episode", d.episode, 4, 1, false}, {"n_success", d.n_success, 4, 1, false},
            {"n_failure", d.n_failure, 4, 1, false}, {"success_steps", d.success_steps, 4, 1, false},
            {"minstd", d.minstd, 4, 1, false}, {"error", d.error, 4, 1, false},
            {"level", d.level, 1, 1, false}, {"check_counter", d.check_counter, 4, 1, false},
            {"win_len", d.win_len, 1, XW_N_T3, false}, {"win_sum", d.win_sum, 1, XW_N_T3, false},
            {"win_pos", d.win_pos, 1, XW_N_T3, false}, {"win_bits", d.win_bits, 4, XW_N_T3 * XW_WIN_WORDS, false}};
        for (auto& t : tbl) if (k == t.n && t.p) { *f = {t.p, t.elem, t.per, t.gm}; return true; }
    } else if (s->cfg.game == XW_GAME_SIMPLE_RACE) {
        XwRaceCfg& r = s->race;
        struct { const char* n; void* p; size_t elem, per; } tbl[] = {
            {"pos_x", r.pos_x, 4, 1}, {"pos_y", r.pos_y, 4, 1}, {"angle", r.angle, 4, 1}, {"steps", r.steps, 4, 1},
            {"state", r.state, 4, 4}, {"error", s->race_error, 4, 1}};
        for (auto& t : tbl) if (k == t.n) { *f = {t.p, t.elem, t.per, false}; return true; }
    }
    return false;
}

static int field_io(xw_sim* s, const char* name, void* h, size_t bytes, bool get) {
    DevGuard dev_guard(s);
    FieldRef f;
    if (!find_field(s, name, &f)) return set_err(XW_ERR_INVALID_ARG, "unknown field '%s'", name);
    const size_t n = s->n;
    CUDA_TRY(cudaDeviceSynchronize());
    if (std::string(name) == "grid") {  // host layout [n][H*W], device rows are CS wide
        const size_t hw = (size_t)s->d.H * s->d.W;
        if (bytes != n * hw) return set_err(XW_ERR_INVALID_ARG, "field grid: expected %zu bytes", n * hw);
        if (get) CUDA_TRY(cudaMemcpy2D(h, hw, f.ptr, s->d.CS, hw, n, cudaMemcpyDeviceToHost));
        else CUDA_TRY(cudaMemcpy2D(f.ptr, s->d.CS, h, hw, hw, n, cudaMemcpyHostToDevice));
        return 0;
    }
    const size_t total = n * f.per_env * f.elem;
    if (bytes != total) return set_err(XW_ERR_INVALID_ARG, "field %s: expected %zu bytes, got %zu", name, total, bytes);
    if (!f.goal_major) {
        if (get) CUDA_TRY(cudaMemcpy(h, f.ptr, total, cudaMemcpyDeviceToHost));
        else CUDA_TRY(cudaMemcpy(f.ptr, h, total, cudaMemcpyHostToDevice));
        return 0;
    }
    // device [goal][env] <-> host [env][goal]
    std::vector<uint8_t> tmp(total);
    if (get) {
        CUDA_TRY(cudaMemcpy(tmp.data(), f.ptr, total, cudaMemcpyDeviceToHost));
        for (size_t g = 0; g < f.per_env; ++g)
            for (size_t e = 0; e < n; ++e) memcpy((uint8_t*)h + (e * f.per_env + g) * f.elem, tmp.data() + (g * n + e) * f.elem, f.elem);
    } else {
        for (size_t g = 0; g < f.per_env; ++g)
            for (size_t e = 0; e < n; ++e) memcpy(tmp.data() + (g * n + e) * f.elem, (const uint8_t*)h + (e * f.per_env + g) * f.elem, f.elem);
        CUDA_TRY(cudaMemcpy(f.ptr, tmp.data(), total, cudaMemcpyHostToDevice));
    }
    return 0;
}

int xw_get_field(xw_sim* s, const char* name, void* h_out, size_t bytes) { return field_io(s, name, h_out, bytes, true); }

int xw_get_fields(xw_sim* s, int32_t n_fields, const char* const* names, void* const* h_out, const size_t* bytes) {
    for (int i = 0; i < n_fields; ++i) {  // (field_io synchronises once: the later calls find an idle device)
        int rc = field_io(s, names[i], h_out[i], bytes[i], true);
        if (rc) return rc;
    }
    return 0;
}

int xw_world_dimensions(const xw_sim* s, double* X, double* Y, double* Z) {
    // XWorldSimulator::get_world_dimensions (xworld_simulator.cpp:100-104); other games leave the values alone
    // (SimulatorInterface::get_world_dimensions only forwards for teaching environments, simulator_interface.cpp:163-167)
    if (s->cfg.game == XW_GAME_XWORLD) { *X = s->cfg.width; *Y = s->cfg.height; *Z = 0; }
    return 0;
}

int xw_extra_info(xw_sim* s, int32_t env, char* buf, size_t cap) {
    if (!buf || cap == 0 || env < 0 || env >= s->n) return set_err(XW_ERR_INVALID_ARG, "xw_extra_info: bad argument");
    if (s->cfg.game != XW_GAME_XWORLD) { buf[0] = 0; return 0; }  // GameSimulator::get_extra_info leaves the string empty
    DevGuard dev_guard(s);
    XwDev& d = s->d;
    CUDA_TRY(cudaDeviceSynchronize());
    uint8_t task = 0, stage = 0, event = 0, level = 0;
    int32_t sit = 0;
    uint32_t minstd = 0;
    CUDA_TRY(cudaMemcpy(&task, d.task + env, 1, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&stage, d.stage + env, 1, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&event, d.event + env, 1, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&sit, d.steps_in_task + env, 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&minstd, d.minstd + env, 4, cudaMemcpyDeviceToHost));
    if (d.level) CUDA_TRY(cudaMemcpy(&level, d.level + env, 1, cudaMemcpyDeviceToHost));
    // the teacher sentence type = the name of the task whose sentence went into the buffer (teaching_task.cpp:118-122): a task
    // records its name whenever no sentence has been recorded yet in this teach -- with walls.json the XWorldRec group comes
    // second and takes the slot unless the navigation task spoke
    const char* type;
    if (d.rules == XW_RULES_NAV3D) type = kT3Names[task < 5 ? task : 0];
    else {
        const bool nav_spoke = (stage == XW_STAGE_NAVIGATION && sit == 0) || event == XW_EVENT_CORRECT_GOAL;
        type = nav_spoke ? kT2Names[task < 4 ? task : 0] : kRecNames[rec_task_of_draw(minstd)];
    }
    static const char* const ev[4] = {"", "correct_goal", "wrong_goal", "time_up"};
    const int side = d.curriculum != 0 ? 3 + level : d.H;  // xworld_.actual_height() = XWorldEnv.get_dims()
    // the reference prints ::getpid() (one process per env); here the env's global id stands in for it
    const int nchar = snprintf(buf, cap, "%lld|task:%s,event:%s,height:%d,width:%d", (long long)(d.gid0 + env), type, ev[event & 3], side, side);
    if (nchar < 0 || (size_t)nchar >= cap) return set_err(XW_ERR_INVALID_ARG, "xw_extra_info: buffer too small");
    return nchar;
}

int xw_task_performance(xw_sim* s, int64_t* successes, int64_t* failures, int64_t* success_steps, int32_t cap, const char** names) {
    if (s->cfg.game != XW_GAME_XWORLD) return 0;
    DevGuard dev_guard(s);
    const int nt = s->d.rules == XW_RULES_NAV3D ? 5 : 4;
    if (cap < nt) return set_err(XW_ERR_INVALID_ARG, "xw_task_performance: room for %d task classes needed", nt);
    unsigned long long h[XW_N_T3 * 3];
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(h, s->d.task_perf, sizeof h, cudaMemcpyDeviceToHost));
    for (int t = 0; t < nt; ++t) {
        successes[t] = (int64_t)h[t * 3]; failures[t] = (int64_t)h[t * 3 + 1]; success_steps[t] = (int64_t)h[t * 3 + 2];
        if (names) names[t] = s->d.rules == XW_RULES_NAV3D ? kT3Names[t] : kT2Names[t];
    }
    return nt;
}

int32_t xw_error_flags(xw_sim* s, int32_t* h_flags) {
    if (s->cfg.game == XW_GAME_SIMPLE_GAME) { if (h_flags) memset(h_flags, 0, sizeof(int32_t) * s->n); return 0; }
    std::vector<int32_t> tmp(s->n);
    if (field_io(s, "error", tmp.data(), sizeof(int32_t) * s->n, true)) return -1;
    int32_t cnt = 0;
    for (int i = 0; i < s->n; ++i) cnt += tmp[i] != 0;
    if (h_flags) memcpy(h_flags, tmp.data(), sizeof(int32_t) * s->n);
    return cnt;
}
int xw_set_field(xw_sim* s, const char* name, const void* h_in, size_t bytes) {
    DevGuard dev_guard(s);
    int rc = field_io(s, name, (void*)h_in, bytes, false);
    if (rc == 0 && s->cfg.game == XW_GAME_XWORLD && s->d.vr > 0 && !strncmp(name, "goal_", 5)) {
        // the cached goal icons are a function of (icon, yaw, scale, offset): warp them again
        k_fpv_warp_goals<<<s->n_sms * 8 < s->n ? s->n_sms * 8 : s->n, 256, 0, s->own_stream>>>(s->d, s->fpv, nullptr, nullptr, nullptr);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(s->own_stream));
    }
    return rc;
}

}  // extern "C"
