"""The reference's process-to-process protocol served from one batch.

In the reference every env is a `SimulatorClient` process that connects to a `SimulatorServer` object in the trainer
(simulator_interface.cpp:165-435) and answers "reset" / "take_actions" / "get_state" / "get_extra_info" / "report_perf"
/ "stop" messages (length-prefixed `util::BinaryBuffer`, packets in `StatePacket::encode` layout).  `BatchClient` plays
N such clients at once: N TCP connections, env i of the batch behind connection i.  The bytes are produced and parsed
by the C ABI (`xw_wire_*`, xworld_b200/csrc/xw_wire.hpp); this file only moves them through sockets and calls the
`Simulator`.

Requests that arrive together are served by ONE batched call: all pending "take_actions" become one step in which the
envs without a request get XW_ACTION_NONE and sit the step out, all pending "reset"s one masked reset.  A trainer that
drives its envs from parallel threads (the reference's design: each SimulatorServer blocks its own thread) therefore
gets full batches; one that drives them one after another still works, one launch per request.
"""
import ctypes as C
import selectors
import socket
import struct

import numpy as np

from . import _abi

U8P = C.POINTER(C.c_uint8)


# ----------------------------------------------------------------------------- packets <-> dicts
def _fields(d):
    """dict -> (XwWireField array, objects to keep alive).  float32 array -> reals, uint8 array / bytes -> pixels,
    int / list of int / int32 array -> ids, str -> str (one part per key, as the reference's games use them)."""
    arr = (_abi.XwWireField * max(1, len(d)))()
    keep = []
    for f, (k, v) in zip(arr, d.items()):
        kb = k.encode()
        keep.append(kb)
        f.key = kb
        if isinstance(v, str):
            sb = v.encode()
            keep.append(sb)
            f.str = sb
            continue
        if isinstance(v, (bytes, bytearray)):
            v = np.frombuffer(bytes(v), np.uint8)
        if isinstance(v, (int, np.integer)):
            v = [int(v)]
        a = np.asarray(v)
        if a.dtype == np.uint8:
            a = np.ascontiguousarray(a).reshape(-1)
            f.pixels, f.n_pixels = a.ctypes.data_as(U8P) if a.size else C.cast(C.c_char_p(b"\0"), U8P), a.size
        elif a.dtype.kind == "f":
            a = np.ascontiguousarray(a, np.float32).reshape(-1)
            f.reals, f.n_reals = (a.ctypes.data_as(C.POINTER(C.c_float)) if a.size
                                  else C.cast(C.c_char_p(b"\0\0\0\0"), C.POINTER(C.c_float))), a.size
        else:
            a = np.ascontiguousarray(a, np.int32).reshape(-1)
            f.ids, f.n_ids = (a.ctypes.data_as(C.POINTER(C.c_int32)) if a.size
                              else C.cast(C.c_char_p(b"\0\0\0\0"), C.POINTER(C.c_int32))), a.size
        keep.append(a)
    return arr, keep


def _sized(call):
    """Run an encoder twice: once to size the buffer, once to fill it."""
    n = call(None, 0)
    if n < 0:
        raise RuntimeError("xw_wire: encode failed (%d)" % n)
    buf = (C.c_uint8 * n)()
    assert call(buf, n) == n
    return bytes(buf)


def _field_value(f):
    if f.reals:
        return np.frombuffer(C.string_at(f.reals, 4 * f.n_reals), np.float32).copy()
    if f.pixels:
        return np.frombuffer(C.string_at(f.pixels, f.n_pixels), np.uint8).copy()
    if f.ids:
        return np.frombuffer(C.string_at(f.ids, 4 * f.n_ids), np.int32).copy()
    return f.str.decode() if f.str is not None else None


def encode_packet(d):
    """StatePacket::encode (data_packet.h:315-321)."""
    lib = _abi.load()
    arr, keep = _fields(d)
    return _sized(lambda out, cap: lib.xw_wire_encode_packet(arr, len(d), out, cap))


def decode_packet(data):
    """StatePacket::decode (data_packet.h:323-333) -> dict."""
    lib = _abi.load()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    arr = (_abi.XwWireField * _abi.XW_WIRE_MAX_FIELDS)()
    n, used = C.c_int32(), C.c_size_t()
    if lib.xw_wire_decode_packet(buf, len(data), arr, _abi.XW_WIRE_MAX_FIELDS, C.byref(n), C.byref(used)) != 0:
        raise RuntimeError("xw_wire: malformed packet")
    return {arr[i].key.decode(): _field_value(arr[i]) for i in range(n.value)}


def parse_request(body):
    """A request body as CommServer::call_remote_func composes it (simulator_communication.h:222-240)."""
    lib = _abi.load()
    buf = (C.c_uint8 * max(1, len(body))).from_buffer_copy(body or b"\0")
    req = _abi.XwWireRequest()
    if lib.xw_wire_parse_request(buf, len(body), C.byref(req)) != 0:
        raise RuntimeError("xw_wire: malformed request")
    return {"cmd": req.cmd.decode(), "act_rep": req.act_rep, "show_screen": bool(req.show_screen), "reward": req.reward,
            "actions": {req.fields[i].key.decode(): _field_value(req.fields[i]) for i in range(req.n_fields)}}


def compose_request(cmd, actions=None, act_rep=1, show_screen=False, reward=0.0):
    """The framed message a SimulatorServer sends (simulator_interface.cpp:184-195,270-299)."""
    lib = _abi.load()
    arr, keep = _fields(actions or {})
    return _sized(lambda out, cap: lib.xw_wire_compose_request(cmd.encode(), arr, len(actions or {}), act_rep, int(show_screen),
                                                               reward, out, cap))


def reply_reset(num_actions, game_over, lives, height, width, channels, X=0.0, Y=0.0, Z=0.0):
    lib = _abi.load()
    return _sized(lambda out, cap: lib.xw_wire_reply_reset(num_actions, game_over, lives, height, width, channels, X, Y, Z, out, cap))


def reply_take_actions(reward, num_steps, game_over, lives, action_success, last_action):
    lib = _abi.load()
    return _sized(lambda out, cap: lib.xw_wire_reply_take_actions(reward, num_steps, game_over, lives, int(action_success),
                                                                  last_action.encode(), out, cap))


def reply_get_state(state):
    lib = _abi.load()
    arr, keep = _fields(state)
    return _sized(lambda out, cap: lib.xw_wire_reply_get_state(arr, len(state), out, cap))


def reply_text(cmd, text=None):
    lib = _abi.load()
    return _sized(lambda out, cap: lib.xw_wire_reply_text(cmd.encode(), None if text is None else text.encode(), out, cap))


# ----------------------------------------------------------------------------- the client loop for a batch
class BatchClient(object):
    """N `SimulatorClient`s (simulator_interface.cpp:316-435) behind one `Simulator` of N envs.

    ports[i] is the port the trainer's i-th SimulatorServer listens on (CommServer::port(), the reference passes it
    to the simulator process on its command line); connection i drives env i."""

    def __init__(self, sim, ports, host="127.0.0.1"):
        if len(ports) != sim.n_envs:
            raise RuntimeError("need one port per env (%d), got %d" % (sim.n_envs, len(ports)))
        self.sim = sim
        self.socks = []
        for p in ports:  # CommClient::establish_connection
            s = socket.create_connection((host, int(p)))
            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            self.socks.append(s)
        self.bufs = [b"" for _ in ports]
        self.open = [True] * len(ports)
        self.errors = []  # (env, reason) of connections that were dropped for a malformed / invalid request
        n = sim.n_envs
        self.reward = np.zeros(n, np.float32)
        self.over = np.zeros(n, np.int32)
        self.success = np.zeros(n, np.uint8)
        self.last_action = [""] * n
        self.steps_served = 0
        self.batches = 0

    # one complete message of connection i, or None
    MAX_FRAME = 1 << 24  # a request is a command word, two scalars and a small action packet; the peer's 64-bit size is not trusted

    def _pop(self, i):
        b = self.bufs[i]
        if len(b) < 8:
            return None
        (size,) = struct.unpack("<Q", b[:8])
        if size > self.MAX_FRAME:
            raise RuntimeError("frame of %d bytes" % size)
        if len(b) < 8 + size:
            return None
        self.bufs[i] = b[8 + size:]
        return b[8:8 + size]

    def _state_of(self, e, reward):
        sim = self.sim
        scr = sim.screen()[e]
        scr = scr.reshape(-1).cpu().numpy() if hasattr(scr, "cpu") else np.asarray(scr).reshape(-1)
        st = {"reward": np.array([reward], np.float32)}
        if sim.cfg.game == _abi.XW_GAME_SIMPLE_RACE:
            st["screen"] = scr.astype(np.float32)
        else:
            st["screen"] = scr.astype(np.uint8)
        if sim.cfg.game == _abi.XW_GAME_XWORLD:  # XWorldSimulator::define_state_specs, xworld_simulator.cpp:486-493
            st["sentence"] = sim.sentences([e])[0]
        return st

    def _extra_info(self, e):  # XWorldSimulator::get_extra_info, xworld_simulator.cpp:495-504 (pid -> env id)
        sim = self.sim
        if sim.cfg.game != _abi.XW_GAME_XWORLD:
            return ""
        return sim.get_extra_info(e)

    def serve(self):
        """SimulatorClient::simulation_loop for every connection, until each has been told to "stop" (or closed)."""
        sim, n = self.sim, self.sim.n_envs
        sel = selectors.DefaultSelector()
        for i, s in enumerate(self.socks):
            sel.register(s, selectors.EVENT_READ, i)
        h, w, c, _ctx = sim.get_screen_out_dimensions()
        if sim.cfg.game == _abi.XW_GAME_SIMPLE_GAME:
            h, c = 1, 1
        n_act = sim.get_num_actions()

        def drop(i, why):
            """One misbehaving connection is closed; the other envs keep being served."""
            self.errors.append((i, why))
            self.open[i] = False
            try:
                sel.unregister(self.socks[i])
            except (KeyError, ValueError):
                pass
            self.socks[i].close()

        def pending():
            """One complete, valid request per open connection, if its buffer already holds one."""
            reqs = {}
            for i in range(n):
                if not self.open[i]:
                    continue
                try:
                    body = self._pop(i)
                    if body is None:
                        continue
                    r = parse_request(body)
                    if r["cmd"] == "take_actions":
                        acts = r["actions"].get("action")
                        if acts is None or len(acts) < 1:
                            raise RuntimeError("take_actions without an 'action'")
                        if r["act_rep"] < 1:
                            raise RuntimeError("act_rep %d" % r["act_rep"])
                        if not 0 <= int(acts[0]) < n_act:
                            raise RuntimeError("action %d outside [0, %d)" % (int(acts[0]), n_act))
                    reqs[i] = (r, body)
                except Exception as ex:  # malformed / oversized / invalid: this connection only
                    drop(i, str(ex))
            return reqs

        while any(self.open):
            reqs = pending()  # requests already buffered (e.g. "stop" sent right behind another message) come first
            if not reqs:
                for key, _ in sel.select():
                    i = key.data
                    try:
                        data = self.socks[i].recv(1 << 20)
                    except OSError as ex:
                        drop(i, str(ex))
                        continue
                    if not data:
                        self.open[i] = False
                        sel.unregister(self.socks[i])
                        continue
                    self.bufs[i] += data
                reqs = pending()
            if not reqs:
                continue
            # ---- one masked reset for every "reset"
            resets = [i for i, (r, _b) in reqs.items() if r["cmd"] == "reset"]
            if resets:
                mask = np.zeros(n, np.uint8)
                mask[resets] = 1
                sim.reset_game(mask)
                self.over[resets] = 0
            # ---- one step for every "take_actions"; the other envs sit it out
            steps = [i for i, (r, _b) in reqs.items() if r["cmd"] == "take_actions"]
            if steps:
                reps = {reqs[i][0]["act_rep"] for i in steps}
                for rep in sorted(reps):  # one launch per distinct act_rep (normally one)
                    a = np.full(n, _abi.XW_ACTION_NONE, np.int32)
                    for i in steps:
                        if reqs[i][0]["act_rep"] == rep:
                            a[i] = int(reqs[i][0]["actions"]["action"][0])
                    r = np.atleast_1d(sim.take_actions(a, rep))
                    o = sim.game_over_codes()
                    m = a != _abi.XW_ACTION_NONE
                    self.reward[m], self.over[m] = r[m], o[m]
                    for i in np.nonzero(m)[0]:
                        self.last_action[i] = str(a[i])
                    self.batches += 1
                if sim.cfg.game == _abi.XW_GAME_XWORLD:
                    self.success[:] = sim.get_field("action_success")
                else:
                    self.success[steps] = 1
                self.steps_served += len(steps)
            num_steps = np.atleast_1d(sim.get_num_steps()) if steps else None
            # ---- replies, in the reference's formats
            for i, (r, body) in reqs.items():
                cmd = r["cmd"]
                lives = 0 if self.over[i] else 1
                if cmd == "reset":
                    X, Y = (float(sim.cfg.width), float(sim.cfg.height)) if sim.cfg.game == _abi.XW_GAME_XWORLD else (0.0, 0.0)
                    out = reply_reset(sim.get_num_actions(), 0, 1, h, w, c, X, Y, 0.0)
                elif cmd == "take_actions":
                    out = reply_take_actions(float(self.reward[i]), int(num_steps[i]), int(self.over[i]), lives,
                                             bool(self.success[i]), self.last_action[i])
                elif cmd == "get_state":
                    out = reply_get_state(self._state_of(i, r["reward"]))
                elif cmd == "get_extra_info":
                    out = reply_text("get_extra_info", self._extra_info(i))
                elif cmd == "stop":  # simulation_loop breaks without answering
                    self.open[i] = False
                    sel.unregister(self.socks[i])
                    self.socks[i].close()
                    continue
                else:  # "report_perf" and anything unknown: deliver_msg sends the received body back
                    out = struct.pack("<Q", len(body)) + body
                try:
                    self.socks[i].sendall(out)
                except OSError as ex:
                    drop(i, str(ex))
        sel.close()
