// xw_wire.hpp -- the reference's process-to-process wire format, host side only.
//
// Replaces (reference file:line), so that a trainer which talks to `SimulatorClient` processes over TCP can attach
// to the envs of one GPU batch instead:
//   util::BinaryBuffer append/read            memory_util.h:83-115,311-386   (native-endian scalars; string = size_t
//                                             length + bytes + NUL; vector<T> = size_t count + elements)
//   StateBuffer::encode/decode                data_packet.cpp:137-174        (u8 flags: 1 reals, 2 pixels, 4 id, 8 str)
//   DataPacket::encode/decode                 data_packet.h:315-333          (size_t n, then key + buffer per entry)
//   MessageHeader / Communicator::deliver_msg simulator_communication.h:34-76, simulator_communication.cpp:31-48
//                                             (a message = size_t body size, then the body)
//   CommServer::call_remote_func              simulator_communication.h:222-240  (body = func name, args, [packet])
//   SimulatorServer requests                  simulator_interface.cpp:184-195,270-299
//   SimulatorClient::simulation_loop, replies simulator_interface.cpp:361-435
//
// No sockets here: these functions turn bytes into requests and replies into bytes; the caller owns the transport
// (xworld_b200/wire.py runs the client loop for a whole batch).  The reference's encoder walks an unordered_map, so
// the ORDER of a packet's entries on the wire is a libstdc++ detail; decoders accept any order, this encoder keeps
// the caller's.
#pragma once
#include <stdint.h>
#include <string.h>

#include "../../include/xworld_b200.h"

namespace xw_wire {

struct Writer {
    uint8_t* out; size_t cap; size_t n;
    void raw(const void* p, size_t k) { if (k && n + k <= cap && out) memcpy(out + n, p, k); n += k; }
    template <typename T> void put(T v) { raw(&v, sizeof v); }
    void str(const char* s) { const size_t l = strlen(s); put<uint64_t>(l); raw(s, l + 1); }
};

struct Reader {
    const uint8_t* in; size_t len; size_t p; bool ok;
    const uint8_t* take(size_t k) { if (!ok || k > len - p) { ok = false; return nullptr; } const uint8_t* q = in + p; p += k; return q; }
    template <typename T> T get() { T v{}; const uint8_t* q = take(sizeof(T)); if (q) memcpy(&v, q, sizeof(T)); return v; }
    const char* str() {  // the NUL is on the wire: the string can be used in place
        const uint64_t l = get<uint64_t>();
        if (!ok || l >= len - p) { ok = false; return nullptr; }
        const char* s = (const char*)take((size_t)l + 1);
        if (!s || s[l] != 0) { ok = false; return nullptr; }
        return s;
    }
};

inline void put_packet(Writer& w, const xw_wire_field* f, int32_t n) {
    w.put<uint64_t>((uint64_t)n);
    for (int32_t i = 0; i < n; ++i) {
        w.str(f[i].key);
        const uint8_t flags = (uint8_t)((f[i].reals ? 1 : 0) | (f[i].pixels ? 2 : 0) | (f[i].ids ? 4 : 0) | (f[i].str ? 8 : 0));
        w.put<uint8_t>(flags);
        if (f[i].reals) { w.put<uint64_t>(f[i].n_reals); w.raw(f[i].reals, sizeof(float) * f[i].n_reals); }
        if (f[i].pixels) { w.put<uint64_t>(f[i].n_pixels); w.raw(f[i].pixels, f[i].n_pixels); }
        if (f[i].ids) { w.put<uint64_t>(f[i].n_ids); w.raw(f[i].ids, sizeof(int32_t) * f[i].n_ids); }
        if (f[i].str) w.str(f[i].str);
    }
}

// An absent vector has a NULL pointer; a present but empty one points at the place in the stream where its
// elements would start (non-NULL, count 0), so encode(decode(x)) == x.
inline bool get_packet(Reader& r, xw_wire_field* f, int32_t max_fields, int32_t* n_fields) {
    const uint64_t n = r.get<uint64_t>();
    if (!r.ok || n > (uint64_t)max_fields) return false;
    for (uint64_t i = 0; i < n; ++i) {
        xw_wire_field& o = f[i];
        memset(&o, 0, sizeof o);
        o.key = r.str();
        const uint8_t flags = r.get<uint8_t>();
        if (!r.ok) return false;
        if (flags & 1) { o.n_reals = r.get<uint64_t>(); if (!r.ok || o.n_reals > (r.len - r.p) / 4) return false; o.reals = (const float*)(r.in + r.p); r.take(4 * (size_t)o.n_reals); }
        if (flags & 2) { o.n_pixels = r.get<uint64_t>(); if (!r.ok || o.n_pixels > r.len - r.p) return false; o.pixels = r.in + r.p; r.take((size_t)o.n_pixels); }
        if (flags & 4) { o.n_ids = r.get<uint64_t>(); if (!r.ok || o.n_ids > (r.len - r.p) / 4) return false; o.ids = (const int32_t*)(r.in + r.p); r.take(4 * (size_t)o.n_ids); }
        if (flags & 8) o.str = r.str();
        if (!r.ok) return false;
    }
    *n_fields = (int32_t)n;
    return true;
}

// deliver_msg: the body's size goes in front of it
inline int64_t frame(Writer& w, size_t cap) {
    const uint64_t body = w.n - 8;
    if (w.out && w.n <= cap) memcpy(w.out, &body, 8);
    return (int64_t)w.n;
}

}  // namespace xw_wire

extern "C" {

int64_t xw_wire_encode_packet(const xw_wire_field* fields, int32_t n_fields, uint8_t* out, size_t cap) {
    if (n_fields < 0 || (n_fields && !fields)) return XW_ERR_INVALID_ARG;
    xw_wire::Writer w{out, cap, 0};
    xw_wire::put_packet(w, fields, n_fields);
    return (int64_t)w.n;
}

int xw_wire_decode_packet(const uint8_t* in, size_t len, xw_wire_field* fields, int32_t max_fields, int32_t* n_fields,
                          size_t* consumed) {
    if (!in || !fields || !n_fields) return XW_ERR_INVALID_ARG;
    xw_wire::Reader r{in, len, 0, true};
    if (!xw_wire::get_packet(r, fields, max_fields, n_fields)) return XW_ERR_INVALID_ARG;
    if (consumed) *consumed = r.p;
    return XW_OK;
}

int xw_wire_parse_request(const uint8_t* body, size_t len, xw_wire_request* req) {
    if (!body || !req) return XW_ERR_INVALID_ARG;
    memset(req, 0, sizeof *req);
    xw_wire::Reader r{body, len, 0, true};
    req->cmd = r.str();
    if (!r.ok) return XW_ERR_INVALID_ARG;
    if (!strcmp(req->cmd, "take_actions")) {  // call_remote_func("take_actions", &actions, act_rep, show_screen)
        req->act_rep = r.get<int32_t>();
        req->show_screen = r.get<uint8_t>();   // bool
        if (!r.ok || !xw_wire::get_packet(r, req->fields, XW_WIRE_MAX_FIELDS, &req->n_fields)) return XW_ERR_INVALID_ARG;
    } else if (!strcmp(req->cmd, "get_state")) {  // call_remote_func("get_state", NULL, reward)
        req->reward = r.get<float>();
        if (!r.ok) return XW_ERR_INVALID_ARG;
    }  // "reset", "report_perf", "get_extra_info", "stop": the name is the whole body
    return XW_OK;
}

int64_t xw_wire_compose_request(const char* cmd, const xw_wire_field* fields, int32_t n_fields, int32_t act_rep, int32_t show_screen,
                        float reward, uint8_t* out, size_t cap) {
    if (!cmd) return XW_ERR_INVALID_ARG;
    xw_wire::Writer w{out, cap, 8};
    w.str(cmd);
    if (!strcmp(cmd, "take_actions")) {
        w.put<int32_t>(act_rep); w.put<uint8_t>(show_screen ? 1 : 0);
        xw_wire::put_packet(w, fields, n_fields);
    } else if (!strcmp(cmd, "get_state")) {
        w.put<float>(reward);
    }
    return xw_wire::frame(w, cap);
}

int64_t xw_wire_reply_reset(int32_t num_actions, int32_t game_over, int32_t lives, uint64_t height, uint64_t width,
                            uint64_t channels, double X, double Y, double Z, uint8_t* out, size_t cap) {
    xw_wire::Writer w{out, cap, 8};
    w.str("reset");
    w.put(num_actions); w.put(game_over); w.put(lives); w.put(height); w.put(width); w.put(channels); w.put(X); w.put(Y); w.put(Z);
    return xw_wire::frame(w, cap);
}

int64_t xw_wire_reply_take_actions(float reward, int64_t num_steps, int32_t game_over, int32_t lives, int32_t action_success,
                                   const char* last_action, uint8_t* out, size_t cap) {
    xw_wire::Writer w{out, cap, 8};
    w.str("take_actions");
    w.put(reward); w.put(num_steps); w.put(game_over); w.put(lives); w.put<uint8_t>(action_success ? 1 : 0);
    w.str(last_action ? last_action : "");
    return xw_wire::frame(w, cap);
}

int64_t xw_wire_reply_get_state(const xw_wire_field* fields, int32_t n_fields, uint8_t* out, size_t cap) {
    if (n_fields < 0 || (n_fields && !fields)) return XW_ERR_INVALID_ARG;
    xw_wire::Writer w{out, cap, 8};
    w.str("get_state");  // compose_msg(state, "get_state"): the name first, the packet last
    xw_wire::put_packet(w, fields, n_fields);
    return xw_wire::frame(w, cap);
}

int64_t xw_wire_reply_text(const char* cmd, const char* text, uint8_t* out, size_t cap) {
    if (!cmd) return XW_ERR_INVALID_ARG;
    xw_wire::Writer w{out, cap, 8};
    w.str(cmd);
    if (text) w.str(text);
    return xw_wire::frame(w, cap);
}

}  // extern "C"
