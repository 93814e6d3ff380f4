// xw_step.cuh -- one SimulatorInterface::take_actions for one env (one warp lane per env).
//
// Replaces (reference file:line):
//   SimulatorInterface::take_actions      simulator_interface.cpp:126-137
//   GameSimulator::take_actions           simulator.cpp:98-108      (num_steps counts calls)
//   XWorldSimulator::take_action          games/xworld/xworld_simulator.cpp:200-265
//   XAgent::act / XMap::move_item         games/xworld/xworld/xitem.cpp:89-155, xmap.cpp:76-101
//   Teacher::teach -> Task::py_stage      teacher.cpp:207-230, teaching_task.cpp:64-116
//   XWorld3DTask._time_reward/_reach_object/_successful_goal/_failed_goal
//                                         games/xworld3d/tasks/xworld3d_task.py:451-482
//   XWorldTask.simple_navigation_reward   games/xworld/tasks/xworld_task.py:184-223
//   XWorldSimulator::game_over + AgentSpecificSimulator::game_over
//                                         xworld_simulator.cpp:165-198, simulator.cpp:158-161
#pragma once
#include "xw_common.cuh"
#include "xw_reset.cuh"

// Rewards are the float32 images of the reference's double sums (Python double -> C++ double
// buffer -> `float r += double`), SURVEY §8a-R.  Written as double expressions folded at compile
// time, so no fp64 instruction reaches the GPU.
#define XW_R3_STEP ((float)(-0.01))
#define XW_R3_CORRECT ((float)(-0.01 + 1.0))
#define XW_R3_WRONG ((float)(-0.01 + -1.0))
#define XW_R2_STEP ((float)(-0.1))
#define XW_R2_STEP_FAILED ((float)(-0.1 + -0.2))

// Batch-wide per-task counters (report only; no state depends on them).
XW_HD void xw_perf_add(const XwDev& d, int task, int which, int v) {
    if (!d.task_perf) return;
#if defined(__CUDA_ARCH__)
    atomicAdd(d.task_perf + task * 3 + which, (unsigned long long)v);
#else
    d.task_perf[task * 3 + which] += (unsigned long long)v;
#endif
}

// Returns true when the env must be reset (auto_reset and the episode ended).
XW_HD bool xw_step_env(const XwDev& d, int e, int action, int act_rep, float* reward_out, int32_t* over_out) {
    const int n = d.n;
    if (action < 0 || action >= (d.vr > 0 ? 6 : 4)) {  // CHECK_LT(action_idx, get_num_actions()) aborts in the reference
        d.error[e] = XW_ERR_INVALID_ACTION;
        if (d.n_invalid) {
#if defined(__CUDA_ARCH__)
            atomicAdd(d.n_invalid, 1);
#else
            ++*d.n_invalid;
#endif
        }
        *reward_out = 0.f;
        *over_out = 0;
        return false;
    }
    uint8_t* g = d.grid + (size_t)e * d.CS;
    int ax = d.agent_x[e], ay = d.agent_y[e];
    int facing = d.facing[e];  // heading codes 0 right 1 down 2 left 3 up (XItem::get_item_facing_dir, xitem.cpp:65-78)
    int dx, dy, move_dir;
    bool turn = false;
    if (d.vr == 0) {
        // MOVE_UP(0,-1) MOVE_DOWN(0,+1) MOVE_LEFT(-1,0) MOVE_RIGHT(+1,0)
        dx = action == 2 ? -1 : action == 3 ? 1 : 0;
        dy = action == 0 ? -1 : action == 1 ? 1 : 0;
        move_dir = action == 0 ? 3 : action == 1 ? 1 : action == 2 ? 2 : 0;
    } else {
        // first-person actions (xitem.cpp:84-86,100-151): MOVE_FORWARD, MOVE_BACKWARD, MOVE_LEFT_FPV, MOVE_RIGHT_FPV relative to
        // the heading; TURN_LEFT / TURN_RIGHT change the yaw by -/+ pi/2 and target the agent's own cell, where
        // XMap::move_item finds the agent itself: not reachable, no contact, returns false (xmap.cpp:76-101) -- a turn is a
        // "failed" action.  The yaw itself is a drifting double in the reference; only its facing class is ever read
        // (get_item_facing_dir's eps 1e-4 against a drift of 1e-16 per turn), so two bits carry it here.
        turn = action >= 4;
        if (turn) {
            for (int rep = 0; rep < act_rep; ++rep) facing = (facing + (action == 4 ? 3 : 1)) & 3;
            d.facing[e] = (uint8_t)facing;
        }
        move_dir = action == 0 ? facing : action == 1 ? (facing + 2) & 3 : action == 2 ? (facing + 3) & 3 : (facing + 1) & 3;
        dx = move_dir == 0 ? 1 : move_dir == 2 ? -1 : 0;
        dy = move_dir == 1 ? 1 : move_dir == 3 ? -1 : 0;
    }
    int num_steps = d.num_steps[e] + 1;
    int collided = XW_CELL_EMPTY, success = 0;
    for (int rep = 0; rep < act_rep && !turn; ++rep) {
        int tx = ax + dx, ty = ay + dy;
        if (tx < 0 || ty < 0 || tx >= d.W || ty >= d.H) { success = 0; break; }  // blocked for all repeats
        int code = g[ty * d.W + tx];
        if (code != XW_CELL_EMPTY) { success = 0; collided = code; break; }
        g[ay * d.W + ax] = XW_CELL_EMPTY;
        g[ty * d.W + tx] = XW_CELL_AGENT;
        ax = tx; ay = ty; success = 1;
    }
    // ---- teacher
    float reward = 0.f;
    int event = XW_EVENT_NONE;
    int stage = d.stage[e];
    if (d.rules == XW_RULES_NAV2D) {
        uint32_t minstd = d.minstd[e];
        if (stage == XW_STAGE_IDLE) {
            int task = d.task[e], tmask = d.tmask[e], aux0 = d.aux0[e];
            int32_t sit = d.steps_in_task[e];
            xw_idle2d(d, d.gid0 + e, (uint32_t)d.episode[e], (uint32_t)num_steps, minstd, d.aux1[e], d.aux2[e],
                      task, stage, tmask, aux0, sit);
            d.task[e] = (uint8_t)task; d.tmask[e] = (uint8_t)tmask; d.aux0[e] = (uint8_t)aux0;
            d.steps_in_task[e] = sit;
        } else {
            reward = success ? XW_R2_STEP : XW_R2_STEP_FAILED;
            const int sit = d.steps_in_task[e] + 1;
            // one_channel only (xworld_task.py:203-210): `steps_in_cur_task >= h*w / 2` (get_max_dims, Python-2 integer
            // division): time up -> _record_failure, back to idle; no event
            if (d.task_mode == XW_TASK_ONE_CHANNEL && sit >= d.H * d.W / 2) {
                d.steps_in_task[e] = 0;
                d.n_failure[e] += 1;
                xw_perf_add(d, d.task[e], 1, 1);
                stage = XW_STAGE_IDLE;
            } else {
                d.steps_in_task[e] = sit;
            }
        }
        xw_minstd_next(minstd);  // XWorldRec group's weighted task sampling: one engine draw per teach
        d.minstd[e] = minstd;
    } else if (stage == XW_STAGE_NAVIGATION) {
        reward = XW_R3_STEP;
        int sit = d.steps_in_task[e] + 1;
        d.steps_in_task[e] = sit;
        // h, w = self.env.get_dims(): the level's world, not the padded map (xworld3d_task.py:475-476)
        const int side = d.curriculum != 0 ? 3 + d.level[e] : d.H;
        if (sit >= side * side * d.max_steps_factor) {
            xw_record_result(d, e, d.task[e], 0);
            d.n_failure[e] += 1;
            xw_perf_add(d, d.task[e], 1, 1);
            event = XW_EVENT_TIME_UP;
            stage = XW_STAGE_TERMINAL;
        } else {
            // _reach_object: |angle(heading, agent->goal)| < pi/4 and goal id in collisions  <=>
            // the blocking item is a goal and the move was along the heading
            const int gr = (collided >= XW_CELL_GOAL0 && move_dir == facing) ? collided - XW_CELL_GOAL0 : -1;
            const int task = d.task[e];
            bool correct = false, wrong = false;
            if (task == XW_T3_BETWEEN) {
                if (gr >= 0) wrong = true;
                else correct = (ax == d.aux1[e] && ay == d.aux2[e]);  // dist(agent, middle) < 0.5
            } else if (task == XW_T3_DIRECTION) {
                if (gr >= 0) {
                    const int ref = d.aux0[e];
                    // the reference's referent is a stale Entity object whose loc cpp_get_entities shifted by the padding
                    // offset in place (xworld_env.py:359-361): at curriculum levels 0-3 the test sees it displaced by
                    // (off, off) -- reproduced, not fixed (oracle/xw_oracle.c xo_teach has the call sequence)
                    const int off = (d.W - side) >> 1;
                    const int rdx = (int)d.goal_x[(size_t)ref * n + e] - (int)d.goal_x[(size_t)gr * n + e] + off;
                    const int rdy = (int)d.goal_y[(size_t)ref * n + e] - (int)d.goal_y[(size_t)gr * n + e] + off;
                    const int fx = facing == 0 ? 1 : facing == 2 ? -1 : 0, fy = facing == 1 ? 1 : facing == 3 ? -1 : 0;
                    const bool close = (rdx * rdx + rdy * rdy) <= 1;  // dist < 1 + 1e-3
                    correct = close && (rdx | rdy) != 0 && xw_axis_direction(fx, fy, rdx, rdy) == d.aux1[e];
                    wrong = !correct;
                }
            } else {  // Target / Near / Avoid: membership in the target set
                if (gr >= 0) { correct = (d.tmask[e] >> gr) & 1; wrong = !correct; }
            }
            if (correct) {
                xw_record_result(d, e, task, 1);
                d.n_success[e] += 1; d.success_steps[e] += sit;
                xw_perf_add(d, task, 0, 1); xw_perf_add(d, task, 2, sit);
                event = XW_EVENT_CORRECT_GOAL; reward = XW_R3_CORRECT; stage = XW_STAGE_TERMINAL;
            } else if (wrong) {
                xw_record_result(d, e, task, 0);
                d.n_failure[e] += 1;
                xw_perf_add(d, task, 1, 1);
                event = XW_EVENT_WRONG_GOAL; reward = XW_R3_WRONG; stage = XW_STAGE_TERMINAL;
            }
        }
    }  // XW_STAGE_TERMINAL: ["terminal", 0, ""]
    int over = 0;
    if (d.max_steps > 0 && num_steps >= d.max_steps) over |= XW_MAX_STEP;
    if (d.task_mode == XW_TASK_LANG_ACQUISITION) {  // one_channel: "all tasks until the max steps" (xworld_simulator.cpp:192-193)
        if (event == XW_EVENT_CORRECT_GOAL) over |= XW_SUCCESS;
        else if (event == XW_EVENT_WRONG_GOAL) over |= XW_DEAD;
        else if (event == XW_EVENT_TIME_UP) over |= XW_MAX_STEP;
    }
    d.agent_x[e] = (uint8_t)ax; d.agent_y[e] = (uint8_t)ay;
    d.num_steps[e] = num_steps;
    d.stage[e] = (uint8_t)stage; d.event[e] = (uint8_t)event; d.succ[e] = (uint8_t)success;
    if (d.ctx_flag) d.ctx_flag[e] = 1;
    *reward_out = reward;
    *over_out = over;
    return d.auto_reset && over != 0;
}
