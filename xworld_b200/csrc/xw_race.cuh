// xw_race.cuh -- SimpleRace step for one env (fp32 state; BASELINE config 5, render off).
//
// Replaces games/simple_race/simple_race_simulator.cpp: BaseCar::move :227-235, RaceEngine::act
// :290-341, get_reward :386-410, get_screen :412-430, reset_game :267-284, StraightTrack /
// CircleTrack geometry :50-101,103-109,182-222, SimpleRaceGame::game_over :465-467.
//
// Numerics: the reference stores floats but its cos/sin/sqrt/fabs resolve to the C double
// functions, so each transcendental is evaluated in fp64 and rounded once to
// fp32; products and sums are single fp32 operations (no FMA contraction: __fmul_rn/__fadd_rn).
#pragma once
#include <math.h>

#include "xw_common.cuh"

#define XW_RACE_PI 3.1415926  // simple_race_simulator.h:39

#if defined(__CUDA_ARCH__)
#define XW_FM(a, b) __fmul_rn((a), (b))
#define XW_FA(a, b) __fadd_rn((a), (b))
#define XW_FD(a, b) __fdiv_rn((a), (b))
#else
#define XW_FM(a, b) ((float)((float)(a) * (float)(b)))
#define XW_FA(a, b) ((float)((float)(a) + (float)(b)))
#define XW_FD(a, b) ((float)((float)(a) / (float)(b)))
#endif

struct XwRaceCfg {
    int32_t n, track_type, full_manouver, difficulty, max_steps, auto_reset;
    float mid_x, mid_y, start_y, end_y, length, width, inner, outer;
    float start_px, start_py;
    double reward_scale;
    float *pos_x, *pos_y, *angle, *state;  // state: [n][4]
    int32_t* steps;
    int32_t random;            // --random (simple_race_simulator.cpp:24): start position and heading drawn per episode
    uint32_t* minstd;          // [n] the env's std::default_random_engine (the reference's per-thread engine, simulator_util.cpp:38-55)
};

// cv::norm(Point2f) returns double: sqrt((double)x*x + (double)y*y)
XW_HD double xw_race_norm(float x, float y) { return sqrt((double)x * (double)x + (double)y * (double)y); }

XW_HD float xw_race_hdisp(const XwRaceCfg& r, float px, float py) {
    if (r.track_type == 0) return XW_FD(XW_FM(2.f, XW_FA(px, -r.mid_x)), r.width);
    return (float)((2 * xw_race_norm(XW_FA(px, -r.mid_x), XW_FA(py, -r.mid_y)) - (double)r.inner - (double)r.outer) / (double)r.width);
}

// util::get_rand_range_val (simulator_util.cpp:57-64): std::uniform_real_distribution<float>(0, upper) on minstd_rand0 as
// libstdc++ evaluates it -- one engine draw, float(x - 1) / 2^31 (the float image of the engine's range 2147483646), clipped
// below 1, times (upper - 0) plus 0 (oracle/xw_oracle.c xo_rand_range_val has the derivation)
XW_HD float xw_rand_range_val(uint32_t& st, float upper) {
    const uint32_t x = xw_minstd_next(st);
    float ret = XW_FD((float)(x - 1u), 2147483648.f);
    if (ret >= 1.0f) ret = 0.99999994f;   // nextafterf(1, 0)
    return XW_FA(XW_FM(ret, upper), 0.0f);
}

// the car of one env (registers; xw_race_load / xw_race_store move it from / to the SoA state)
struct XwRaceCar { float px, py, ang; int32_t steps; uint32_t minstd; };
XW_HD XwRaceCar xw_race_load(const XwRaceCfg& r, int e) {
    XwRaceCar c;
    c.px = r.pos_x[e]; c.py = r.pos_y[e]; c.ang = r.angle[e]; c.steps = r.steps[e]; c.minstd = r.random ? r.minstd[e] : 0u;
    return c;
}
XW_HD void xw_race_store(const XwRaceCfg& r, int e, const XwRaceCar& c) {
    r.pos_x[e] = c.px; r.pos_y[e] = c.py; r.angle[e] = c.ang; r.steps[e] = c.steps;
    if (r.random) r.minstd[e] = c.minstd;
}

XW_HD void xw_race_reset_car(const XwRaceCfg& r, XwRaceCar& c) {
    c.steps = 0;
    if (r.random) {  // RaceEngine::reset_game :267-284 with --random: track draw, start position, heading (four draws)
        uint32_t st = c.minstd;
        (void)xw_rand_range_val(st, 1.0f);   // the track index: one track in the pool, the draw is still taken
        if (r.track_type == 0) {   // StraightTrack::get_start_pos :192-199
            const float dy = XW_FD(XW_FM(xw_rand_range_val(st, 1.0f), r.length), 2.f);
            const float dx = (float)(((double)xw_rand_range_val(st, 1.0f) - 0.5) * (double)r.width);
            c.px = XW_FA(dx, r.start_px); c.py = XW_FA(dy, r.start_py);
        } else {                   // CircleTrack::get_start_pos :78-86
            const float theta = (float)((double)XW_FM(xw_rand_range_val(st, 1.0f), 2.f) * XW_RACE_PI);
            const float rad = XW_FA(r.inner, XW_FM(xw_rand_range_val(st, 1.0f), r.width));
            c.px = XW_FA((float)((double)rad * cos((double)theta)), r.mid_x);
            c.py = XW_FA((float)((double)rad * sin((double)theta)), r.mid_y);
        }
        c.ang = (float)((double)XW_FM(xw_rand_range_val(st, 1.0f), 2.f) * XW_RACE_PI);   // BaseCar::set_angle :237-243
        c.minstd = st;
        return;
    }
    c.px = r.start_px; c.py = r.start_py;
    c.ang = (float)(XW_RACE_PI / 2);
}
XW_HD void xw_race_reset_env(const XwRaceCfg& r, int e) {
    XwRaceCar c = xw_race_load(r, e);
    xw_race_reset_car(r, c);
    xw_race_store(r, e, c);
}

// One GameSimulator::take_actions (simulator.cpp:98-108): num_steps counts the call, the action is applied act_rep times
// (RaceEngine::act moves the car whether or not it has left the track) and the float rewards are summed; the state vector
// and game_over describe the car after the last repeat.
XW_HD bool xw_race_step_car(const XwRaceCfg& r, XwRaceCar& c, int action_index, int act_rep, float* reward_out, int32_t* over_out, float* st) {
    int a = r.full_manouver ? action_index : (action_index == 0 ? 4 : 7);
    const float delta_ang = (float)(XW_RACE_PI / 10), delta_fwd = 1.f;
    float d_forward = 0.f, d_turn = 0.f;
    int m = a % 3;
    if (m == 1) d_forward = delta_fwd; else if (m == 2) d_forward = -delta_fwd;
    m = (a / 3) % 3;
    if (m == 1) d_turn = delta_ang; else if (m == 2) d_turn = -delta_ang;
    float ang = c.ang, px = c.px, py = c.py;
    const int steps = c.steps + 1;
    float total = 0.f, tx = 0.f, ty = 1.f, hd = 0.f;
    double ca = 0, sa = 0;
    bool oob = false;
    for (int rep = 0; rep < act_rep; ++rep) {
        ang = XW_FA(ang, d_turn);
        if ((double)ang > 2 * XW_RACE_PI) ang = (float)((double)ang - 2 * XW_RACE_PI);
        else if (ang < 0) ang = (float)((double)ang + 2 * XW_RACE_PI);
#if defined(__CUDA_ARCH__)
        sincos((double)ang, &sa, &ca);  // one argument reduction for both (same polynomials as cos() / sin())
#else
        ca = cos((double)ang); sa = sin((double)ang);
#endif
        const float cx = (float)ca, sx = (float)sa;
        px = XW_FA(px, XW_FM(d_forward, cx));
        py = XW_FA(py, XW_FM(d_forward, sx));
        // tangent
        if (r.track_type == 0) { tx = 0.f; ty = 1.f; }
        else {
            float ax = XW_FA(r.mid_y, -py), ay = XW_FA(px, -r.mid_x);
            double s = 1 / xw_race_norm(ax, ay);
            tx = (float)((double)ax * s); ty = (float)((double)ay * s);
        }
        const float reward_speed = XW_FM(XW_FA(XW_FM(cx, tx), XW_FM(sx, ty)), d_forward);
        const bool finish = (r.track_type == 0) && (py > r.end_y);
        if (r.track_type == 0) {
            const float hw = XW_FD(r.width, 2.f);
            oob = (px < XW_FA(r.mid_x, -hw)) || (px > XW_FA(r.mid_x, hw)) || (py < r.start_y) || (py > r.end_y);
        } else {
            float rr = (float)xw_race_norm(XW_FA(px, -r.mid_x), XW_FA(py, -r.mid_y));
            oob = rr < r.inner || rr > r.outer;
        }
        hd = xw_race_hdisp(r, px, py);
        const float reward_finish = finish ? 2.f : 0.f;
        const float reward_boundary = r.difficulty == 0 ? (float)(-fabs((double)hd)) : ((oob && !finish) ? -2.f : 0.f);
        float reward = XW_FA(XW_FA(reward_finish, reward_boundary), reward_speed);
        reward = (float)((double)reward * r.reward_scale);
        total = rep == 0 ? reward : XW_FA(total, reward);   // (0.f + x == x, also for x == -0.f + 0.f ... kept explicit)
    }
    // state (get_screen)
    double ct = (double)tx * ca + (double)ty * sa;
    ct = ct < -1.0 ? -1.0 : (ct > 1.0 ? 1.0 : ct);
    const float cos_theta = (float)ct;
    float sin_theta = (float)sqrt((double)XW_FA(1.f, -XW_FM(cos_theta, cos_theta)));
    if (ca * (double)ty + sa * (double)tx < 0) sin_theta = -sin_theta;
    st[0] = cos_theta; st[1] = sin_theta; st[2] = hd;
    st[3] = r.track_type == 0 ? XW_FD(XW_FM(2.f, XW_FA(py, -r.mid_y)), r.length) : 0.f;
    int over = 0;
    if (r.max_steps > 0 && steps >= r.max_steps) over |= XW_MAX_STEP;
    if (oob) over |= XW_DEAD;
    c.px = px; c.py = py; c.ang = ang; c.steps = steps;
    *reward_out = total;
    *over_out = over;
    return r.auto_reset && over != 0;
}
XW_HD bool xw_race_step_env(const XwRaceCfg& r, int e, int action_index, int act_rep, float* reward_out, int32_t* over_out) {
    XwRaceCar c = xw_race_load(r, e);
    const bool need = xw_race_step_car(r, c, action_index, act_rep, reward_out, over_out, r.state + (size_t)e * 4);
    xw_race_store(r, e, c);
    return need;
}
