// xworld_b200.hpp -- header-only C++ host side over the C ABI (xworld_b200.h).
//
// Mirrors the reference's in-process interface class for this path, method for method:
//   simulator::SimulatorInterface            /root/reference simulator_interface.h:40-89
//     ctor by game name                      simulator_interface.cpp:37-85
//     reset_game                             simulator_interface.cpp:95-105
//     game_over / game_over_string           simulator_interface.cpp:107-113, simulator.cpp:125-144
//     take_actions / take_action             simulator_interface.cpp:126-137, simulator_interface.h:66-68
//     get_state                              simulator_interface.cpp:139-143 (screen + reward of StatePacket)
//     get_num_actions / get_lives / get_num_steps / get_screen_out_dimensions / last_action_success
//     get_extra_info / teacher_report_task_performance / last_action / get_world_dimensions   simulator_interface.h:72-80
// with the same argument meaning and, where the reference aborts (CHECK / LOG(FATAL)), a
// std::runtime_error instead.  The reference class owns ONE environment; this one owns a batch of
// n_envs (n_envs = 1 reproduces the reference's call shapes through the scalar overloads).
// Process-global gflags become the xw_config passed to the constructor.
#ifndef XWORLD_B200_HPP_
#define XWORLD_B200_HPP_

#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "xworld_b200.h"

namespace xworld_b200 {

// GameSimulator::decode_game_over_code (simulator.cpp:125-144)
inline std::string decode_game_over_code(int code) {
    if (code == XW_ALIVE) return "alive";
    std::string s;
    const struct { int bit; const char* name; } names[] = {
        {XW_MAX_STEP, "max_step"}, {XW_DEAD, "dead"}, {XW_SUCCESS, "success"}, {XW_LOST_LIFE, "lost_life"}};
    for (const auto& n : names)
        if (code & n.bit) { if (!s.empty()) s += "|"; s += n.name; }
    return s;
}

// What the reference returns as StatePacket{"reward", "screen"} (data_packet.h:347-379): the bytes of
// get_state()["screen"] (context*C*H*W per env, planes B,G,R) and the reward that was passed in.
struct State {
    std::vector<float> reward;    // [n_envs]
    std::vector<uint8_t> screen;  // [n_envs][context*C][H][W]; simple_race: raw float bytes of the 4-float state
};

class SimulatorInterface {
public:
    // name: "xworld" | "simple_game" | "simple_race" (simulator_interface.cpp:37-85).  cfg.game is set
    // from the name; catalog is required for "xworld" only.
    SimulatorInterface(const std::string& name, xw_config cfg, const xw_catalog* catalog = nullptr, int n_envs = 1,
                       int device = -1)
        : running_(false), n_(n_envs) {
        if (name == "xworld") cfg.game = XW_GAME_XWORLD;
        else if (name == "simple_game") cfg.game = XW_GAME_SIMPLE_GAME;
        else if (name == "simple_race") cfg.game = XW_GAME_SIMPLE_RACE;
        else throw std::runtime_error("Unrecognized game type: " + name);
        check(xw_create(&cfg, catalog, n_envs, device, &sim_));
        int32_t h, w, c, k;
        xw_screen_dims(sim_, &h, &w, &c, &k);
        h_ = h; w_ = w; c_ = c; context_ = k;
        frame_bytes_ = xw_frame_bytes(sim_);
        screen_.assign((size_t)n_ * frame_bytes_, 0);
        over_.assign(n_, 0);
        reward_.assign(n_, 0.f);
        acc_reward_.assign(n_, 0.f);
    }
    SimulatorInterface(const SimulatorInterface&) = delete;
    SimulatorInterface& operator=(const SimulatorInterface&) = delete;
    virtual ~SimulatorInterface() { xw_destroy(sim_); }

    static xw_config default_config() { xw_config c; xw_config_init(&c); return c; }

    virtual void start() { running_ = true; }
    virtual void stop() { running_ = false; }

    // reset every env (mask == nullptr) or the envs with mask[i] != 0
    virtual void reset_game(const uint8_t* mask = nullptr) {
        check(xw_reset_host(sim_, mask, screen_.data()));
        for (int i = 0; i < n_; ++i)
            if (!mask || mask[i]) { over_[i] = 0; acc_reward_[i] = 0.f; }
    }

    virtual int game_over(int env = 0) const { return over_.at(env); }
    virtual std::string game_over_string(int env = 0) const { return decode_game_over_code(over_.at(env)); }
    virtual int get_num_actions() const { return xw_num_actions(sim_); }
    virtual int get_lives(int env = 0) const { return over_.at(env) ? 0 : 1; }  // xworld_simulator.cpp:506
    virtual int64_t get_num_steps(int env = 0) {
        std::vector<int64_t> v(n_);
        check(xw_num_steps(sim_, v.data()));
        return v.at(env);
    }
    virtual void get_screen_out_dimensions(size_t& height, size_t& width, size_t& channels) const {
        height = h_; width = w_; channels = c_;
    }
    int context() const { return context_; }
    int num_envs() const { return n_; }

    // take_actions for the whole batch: actions[n_envs] = the "action" id of each env's StatePacket.
    // Returns the rewards ([n_envs], valid until the next call).  show_screen is the reference's GUI
    // switch and must be false here.
    virtual const std::vector<float>& take_actions(const std::vector<int32_t>& actions, int act_rep = 1,
                                                   bool show_screen = false) {
        if (show_screen) throw std::runtime_error("show_screen is not supported (no GUI)");
        if ((int)actions.size() != n_) throw std::runtime_error("expected one action per env");
        std::vector<float> r(n_);
        std::vector<int32_t> o(n_);
        const int rc = xw_step_host(sim_, actions.data(), act_rep, r.data(), o.data(), screen_.data());
        if (rc != XW_OK && rc != XW_ERR_INVALID_ACTION) check(rc);
        last_action_.resize(n_);
        for (int i = 0; i < n_; ++i)
            if (actions[i] != XW_ACTION_NONE) {  // an env that sat the step out keeps its totals, reward and game_over
                reward_[i] = r[i]; over_[i] = o[i]; acc_reward_[i] += r[i];
                last_action_[i] = std::to_string(actions[i]);  // XWorldSimulator::take_action: last_action_ (xworld_simulator.cpp:203,258)
            }
        if (rc == XW_ERR_INVALID_ACTION) check(rc);  // the valid envs were stepped; the reference CHECK-aborts here
        return reward_;
    }
    // the reference's scalar shape (n_envs == 1)
    virtual float take_actions(int action, int act_rep, bool show_screen) {
        return take_actions(std::vector<int32_t>((size_t)n_, action), act_rep, show_screen)[0];
    }
    float take_action(int action, bool show_screen = false) { return take_actions(action, 1, show_screen); }

    // get_state(reward): a deep copy of the context frames + the reward passed in (simulator.cpp:87-96)
    virtual State get_state(float reward = 0.f) const {
        State s;
        s.reward.assign(n_, reward);
        s.screen = screen_;
        return s;
    }
    const std::vector<uint8_t>& screen() const { return screen_; }
    size_t frame_bytes() const { return frame_bytes_; }

    virtual bool last_action_success(int env = 0) {
        std::vector<uint8_t> v(n_);
        check(xw_get_field(sim_, "action_success", v.data(), v.size()));
        return v.at(env) != 0;
    }
    // GameSimulator::last_action (the action id as text; xworld_simulator.cpp:203,258)
    virtual std::string last_action(int env = 0) const { return env < (int)last_action_.size() ? last_action_[env] : std::string(); }
    // XWorldSimulator::get_extra_info (xworld_simulator.cpp:495-504): "<id>|task:..,event:..,height:..,width:.."; "" for other games
    virtual void get_extra_info(std::string& info, int env = 0) {
        char buf[256];
        const int rc = xw_extra_info(sim_, env, buf, sizeof buf);
        if (rc < 0) check(rc);
        info.assign(buf);
    }
    // SimulatorInterface::get_world_dimensions (simulator_interface.cpp:163-167): untouched unless the game is a teaching environment
    virtual void get_world_dimensions(double& X, double& Y, double& Z) const { check(xw_world_dimensions(sim_, &X, &Y, &Z)); }
    // Teacher::report_task_performance (teacher.cpp:175-200): the lines the reference logs, summed over the batch
    virtual std::vector<std::string> teacher_report_task_performance() {
        int64_t s[8], f[8], st[8];
        const char* names[8];
        const int nt = xw_task_performance(sim_, s, f, st, 8, names);
        if (nt < 0) check(nt);
        std::vector<std::string> lines;
        for (int t = 0; t < nt; ++t) {
            lines.push_back(std::string("=== ") + names[t] + " ===");
            if (s[t] + f[t] == 0) continue;  // "skip task that did not occur"
            char buf[160];
            const double per = s[t] > 0 ? (double)st[t] / (double)s[t] : -1;
            snprintf(buf, sizeof buf, "=== %lld(S)/%lld(F) -> %g@%g", (long long)s[t], (long long)f[t], (double)s[t] / (double)(s[t] + f[t]), per);
            lines.push_back(buf);
        }
        return lines;
    }
    float acc_reward(int env = 0) const { return acc_reward_.at(env); }
    xw_sim* handle() { return sim_; }  // for the device-pointer entry points (xw_step / xw_render)

protected:
    static void check(int rc) {
        if (rc != XW_OK) throw std::runtime_error(std::string("xworld_b200: ") + xw_last_error());
    }
    bool running_;
    int n_;
    xw_sim* sim_ = nullptr;
    size_t h_ = 0, w_ = 0, c_ = 0, frame_bytes_ = 0;
    int context_ = 1;
    std::vector<uint8_t> screen_;
    std::vector<int32_t> over_;
    std::vector<float> reward_, acc_reward_;
    std::vector<std::string> last_action_;
};

}  // namespace xworld_b200
#endif  // XWORLD_B200_HPP_
