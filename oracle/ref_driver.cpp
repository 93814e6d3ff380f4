// ref_driver.cpp -- C entry points over the REFERENCE's own translation units, compiled unmodified
// from /root/reference against oracle/shim (see oracle/Makefile).  TEST INFRASTRUCTURE ONLY.
//
// Exposes: SimpleGame (games/simple_game), SimpleRaceGame physics (games/simple_race), XMap/XAgent
// step logic (games/xworld/xworld/xmap.cpp, xitem.cpp) and util::get_rand_ind (simulator_util.cpp).
#include <algorithm>
#include <cstring>
#include <string>
#include <memory>
#include <thread>
#include <vector>

#include "games/simple_game/simple_game_simulator.h"
#include "games/simple_race/simple_race_simulator.h"
#include "games/xworld/xworld/xmap.h"
#include "simulator_util.h"
#include "memory_util.h"
#include "data_packet.h"

DECLARE_int32(array_size);
DECLARE_int32(simulator_seed);
DECLARE_int32(max_steps);
DECLARE_string(track_type);
DECLARE_double(track_width);
DECLARE_double(track_length);
DECLARE_double(track_radius);
DECLARE_bool(race_full_manouver);
DECLARE_bool(random);
DECLARE_string(difficulty);
DECLARE_double(reward_scale);
DEFINE_int32(visible_radius, 0, "xworld_simulator.cpp:26 (that TU is not compiled here)");

using namespace simulator;

extern "C" {

// ---- SimpleGame through the GameSimulator interface (tests/test_simple_game_simulator.cpp) ----
void* ref_sg_create(int array_size) { FLAGS_array_size = array_size; return new simple_game::SimpleGame(); }
void ref_sg_destroy(void* p) { delete (simple_game::SimpleGame*)p; }
void ref_sg_reset(void* p) { ((simple_game::SimpleGame*)p)->reset_game(); }
float ref_sg_take_action(void* p, int a) {
    StatePacket act;
    act.add_buffer_id("action", {a});
    return ((simple_game::SimpleGame*)p)->take_action(act);
}
int ref_sg_game_over(void* p) { return ((simple_game::SimpleGame*)p)->game_over(); }
void ref_sg_screen(void* p, uint8_t* out, int n) {
    StatePacket s;
    ((simple_game::SimpleGame*)p)->get_screen(s);
    std::memcpy(out, s.get_buffer("screen")->get_value<uint8_t>(), n);
}

// ---- SimpleRace ----
// --random: RaceEngine reads FLAGS_random in its constructor (:257-258); the start state is drawn from the calling thread's
// engine (util::get_rand_range_val), re-seeded here the way std::default_random_engine::seed does
static bool g_race_random = false;
void ref_race_set_random(int on) { g_race_random = on != 0; }   // applies to the next ref_race_create
void ref_seed_thread_engine(unsigned seed) { util::thread_local_reng().seed(seed); }
void* ref_race_create(int track_type, double width, double length, double radius, int full, int hard, double scale, int max_steps) {
    FLAGS_track_type = track_type == 0 ? "straight" : "circle";
    FLAGS_track_width = width; FLAGS_track_length = length; FLAGS_track_radius = radius;
    FLAGS_race_full_manouver = full != 0; FLAGS_random = g_race_random;
    FLAGS_difficulty = hard ? "hard" : "easy"; FLAGS_reward_scale = scale; FLAGS_max_steps = max_steps;
    return new simple_race::SimpleRaceGame();
}
void ref_race_destroy(void* p) { delete (simple_race::SimpleRaceGame*)p; }
void ref_race_reset(void* p) { ((simple_race::SimpleRaceGame*)p)->reset_game(); }
float ref_race_step(void* p, int action_index, float* state4, int* game_over) {
    auto* g = (simple_race::SimpleRaceGame*)p;
    StatePacket act;
    act.add_buffer_id("action", {action_index});
    float r = g->take_action(act);
    StatePacket s;
    g->get_screen(s);
    std::memcpy(state4, s.get_buffer("screen")->get_value<float>(), 4 * sizeof(float));
    *game_over = g->game_over();
    return r;
}

// the reference's own GameSimulator::take_actions (simulator.cpp:98-108) around SimpleRaceGame::take_action
float ref_race_take_actions(void* p, int action_index, int act_rep, float* state4, int* game_over) {
    auto* g = (simple_race::SimpleRaceGame*)p;
    StatePacket act;
    act.add_buffer_id("action", {action_index});
    float r = g->take_actions(act, act_rep, false, 0.0f);
    StatePacket s;
    g->get_screen(s);
    std::memcpy(state4, s.get_buffer("screen")->get_value<float>(), 4 * sizeof(float));
    *game_over = g->game_over();
    return r;
}

// ---- XMap + XAgent (xmap.cpp:76-101, xitem.cpp:89-155) ----
struct RefMap {
    xwd::XMap map;
    std::vector<xwd::XItemPtr> items;
    xwd::XItemPtr agent;
};
// types[i]: 0 block, 1 goal, 2 agent; ids are "e<i>"
void* ref_map_create(int h, int w, int n, const int* types, const int* xs, const int* ys, double agent_yaw, int visible_radius) {
    FLAGS_visible_radius = visible_radius;
    RefMap* m = new RefMap();
    m->map = xwd::XMap(h, w);
    for (int i = 0; i < n; ++i) {
        Entity e;
        e.type = types[i] == 0 ? "block" : types[i] == 1 ? "goal" : "agent";
        e.id = "e" + std::to_string(i);
        e.loc = Vec3(xs[i], ys[i], 0);
        e.yaw = types[i] == 2 ? agent_yaw : 1.5707963;
        e.scale = 1.0; e.offset = 0.0;
        e.name = e.type; e.asset_path = ""; e.color = "na";
        auto it = xwd::XItem::create_item(e);
        m->items.push_back(it);
        if (types[i] == 2) m->agent = it;
    }
    m->map.add_items(m->items);
    return m;
}
void ref_map_destroy(void* p) { delete (RefMap*)p; }
// XWorld::act (xworld.cpp:162-166).  contact = index of the first contacted entity or -1.
int ref_map_act(void* p, int action, int* ax, int* ay, double* yaw, int* contact, int* n_contacts) {
    RefMap* m = (RefMap*)p;
    std::vector<std::string> contacts;
    xwd::Loc target = m->agent->act(action);
    bool ok = m->map.move_item(m->agent, target, contacts);
    xwd::Loc l = m->agent->get_item_location();
    *ax = l.x; *ay = l.y; *yaw = m->agent->get_item_yaw();
    *n_contacts = (int)contacts.size();
    *contact = contacts.empty() ? -1 : std::stoi(contacts[0].substr(1));
    return ok ? 1 : 0;
}
// XMap::image_masking (xmap.cpp:273-362) at the agent's location and yaw: rect4 = {x, y, width, height} in cells of the
// padded map, shadow = vr*vr flags
void ref_map_masking(void* p, int vr, int* rect4, uint8_t* shadow) {
    RefMap* m = (RefMap*)p;
    std::vector<bool> sh;
    cv::Rect r = m->map.image_masking(m->agent->get_item_location(), m->agent->get_item_yaw(), vr, sh);
    rect4[0] = r.x; rect4[1] = r.y; rect4[2] = r.width; rect4[3] = r.height;
    for (size_t i = 0; i < sh.size(); ++i) shadow[i] = sh[i] ? 1 : 0;
}
int ref_map_num_actions(void* p) { return ((RefMap*)p)->agent->get_num_actions(); }

// ---- wire format: the reference's own util::BinaryBuffer (memory_util.h) and StatePacket::encode/decode
// (data_packet.h:315-333, data_packet.cpp:137-174).  The message bodies are composed with the same append sequence
// as CommServer::call_remote_func / Communicator::compose_msg (simulator_communication.h:160-176,222-240) and the
// SimulatorClient replies (simulator_interface.cpp:385-435); deliver_msg's header insert is simulator_communication.cpp:31-38
// (that TU needs boost::asio and is not compiled here).
struct RefField { const char* key; const float* reals; uint64_t n_reals; const uint8_t* pixels; uint64_t n_pixels;
                  const int* ids; uint64_t n_ids; const char* str; };
static void ref_fill_packet(StatePacket& p, const RefField* f, int n) {
    for (int i = 0; i < n; ++i) {
        p.add_key(f[i].key);
        auto b = p.get_buffer(f[i].key);
        if (f[i].reals) b->set_value(f[i].reals, f[i].reals + f[i].n_reals);
        if (f[i].pixels) b->set_value(f[i].pixels, f[i].pixels + f[i].n_pixels);
        if (f[i].ids) b->set_id(f[i].ids, f[i].ids + f[i].n_ids);
        if (f[i].str) b->set_str(f[i].str);
    }
}
static long ref_emit(util::BinaryBuffer& buf, bool framed, uint8_t* out, long cap) {
    if (framed) buf.insert(0, (size_t)buf.size());  // MessageHeader::insert_into_msg
    if ((long)buf.size() <= cap) std::memcpy(out, buf.data(), buf.size());
    return (long)buf.size();
}
long ref_wire_encode_packet(const RefField* f, int n, uint8_t* out, long cap) {
    StatePacket p; ref_fill_packet(p, f, n);
    util::BinaryBuffer buf; p.encode(buf);
    return ref_emit(buf, false, out, cap);
}
// decodes with the reference's code and writes a canonical text dump (keys sorted) for comparison
long ref_wire_decode_dump(const uint8_t* in, long len, char* out, long cap) {
    util::BinaryBuffer buf(in, (size_t)len);
    buf.rewind();
    StatePacket p; p.decode(buf);
    auto keys = p.get_keys();
    std::sort(keys.begin(), keys.end());
    std::string s;
    for (auto& k : keys) {
        auto b = p.get_buffer(k);
        s += k + "|";
        util::BinaryBuffer one; b->encode(one);   // flags + present parts, the reference's own layout
        static const char* hex = "0123456789abcdef";
        for (size_t i = 0; i < one.size(); ++i) { s += hex[one.data()[i] >> 4]; s += hex[one.data()[i] & 15]; }
        s += "\n";
    }
    if ((long)s.size() + 1 <= cap) std::memcpy(out, s.c_str(), s.size() + 1);
    return (long)s.size() + 1;
}
long ref_wire_request(const char* cmd, const RefField* f, int n, int act_rep, int show_screen, float reward, uint8_t* out, long cap) {
    util::BinaryBuffer buf;
    std::string name(cmd);
    buf.append(name);
    if (name == "take_actions") {   // compose_msg(*sim_data, func_name, act_rep, show_screen): args first, packet last
        buf.append(act_rep); buf.append((bool)show_screen);
        StatePacket p; ref_fill_packet(p, f, n); p.encode(buf);
    } else if (name == "get_state") {
        buf.append(reward);
    }
    return ref_emit(buf, true, out, cap);
}
long ref_wire_reply_reset(int num_actions, int game_over, int lives, size_t h, size_t w, size_t c, double X, double Y, double Z,
                          uint8_t* out, long cap) {
    util::BinaryBuffer buf;
    buf.append(std::string("reset")); buf.append(num_actions); buf.append(game_over); buf.append(lives);
    buf.append(h); buf.append(w); buf.append(c); buf.append(X); buf.append(Y); buf.append(Z);
    return ref_emit(buf, true, out, cap);
}
long ref_wire_reply_take_actions(float reward, int64_t num_steps, int game_over, int lives, int success, const char* last_action,
                                 uint8_t* out, long cap) {
    util::BinaryBuffer buf;
    buf.append(std::string("take_actions")); buf.append(reward); buf.append(num_steps); buf.append(game_over); buf.append(lives);
    buf.append((bool)success); buf.append(std::string(last_action));
    return ref_emit(buf, true, out, cap);
}
long ref_wire_reply_get_state(const RefField* f, int n, uint8_t* out, long cap) {
    util::BinaryBuffer buf;
    buf.append(std::string("get_state"));
    StatePacket p; ref_fill_packet(p, f, n); p.encode(buf);
    return ref_emit(buf, true, out, cap);
}
long ref_wire_reply_text(const char* cmd, const char* text, uint8_t* out, long cap) {
    util::BinaryBuffer buf;
    buf.append(std::string(cmd));
    if (text) buf.append(std::string(text));
    return ref_emit(buf, true, out, cap);
}
// reads a take_actions reply the way SimulatorServer::take_actions does (simulator_interface.cpp:279-282)
int ref_wire_read_take_actions_reply(const uint8_t* body, long len, float* r, int64_t* num_steps, int* game_over, int* lives,
                                     int* success, char* last_action, long cap) {
    util::BinaryBuffer buf(body, (size_t)len);
    buf.rewind();
    std::string reply, la; bool ok;
    buf.read(reply); buf.read(*r); buf.read(*num_steps); buf.read(*game_over); buf.read(*lives); buf.read(ok); buf.read(la);
    *success = ok;
    if ((long)la.size() + 1 <= cap) std::memcpy(last_action, la.c_str(), la.size() + 1);
    return reply == "take_actions" && buf.eof() ? 0 : -1;
}

// ---- util::simple_importance_sampling (simulator_util.cpp:75-86) on a fresh thread: `n_draws` consecutive samples over the
// accumulated weights, and the engine's raw output that each sample consumed (the thread's engine is re-seeded the same way on
// a second fresh thread and stepped alongside)
void ref_importance_sampling(int simulator_seed, const double* acc_weights, int n_weights, int n_draws, int* out_idx) {
    FLAGS_simulator_seed = simulator_seed;
    std::vector<double> acc(acc_weights, acc_weights + n_weights);
    std::thread th([&]() { for (int i = 0; i < n_draws; ++i) out_idx[i] = util::simple_importance_sampling(acc); });
    th.join();
}

// ---- util::get_rand_ind on fresh threads (tests/test_simulator_seed.cpp) ----
void ref_rand_ind_threads(int simulator_seed, int n_threads, int size, int* out) {
    FLAGS_simulator_seed = simulator_seed;
    for (int i = 0; i < n_threads; ++i) {
        std::thread th([&, i]() { out[i] = util::get_rand_ind(size); });
        th.join();
    }
}
}
