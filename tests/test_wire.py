"""the SimulatorServer/Client wire format.  The C ABI's encoders / parsers (xw_wire_*) against the
reference's own util::BinaryBuffer + StatePacket code compiled in oracle/_ref (ref_driver.cpp ref_wire_*), byte for
byte, and the batch client loop over real TCP sockets with the trainer side played by the reference's composer."""
import ctypes as C
import os
import socket
import struct
import threading

import numpy as np
import pytest

import oracle
from xworld_b200 import _abi, wire
from xworld_b200.simulator import Simulator


class RefField(C.Structure):
    _fields_ = [("key", C.c_char_p), ("reals", C.POINTER(C.c_float)), ("n_reals", C.c_uint64),
                ("pixels", C.POINTER(C.c_uint8)), ("n_pixels", C.c_uint64), ("ids", C.POINTER(C.c_int)), ("n_ids", C.c_uint64),
                ("str", C.c_char_p)]


@pytest.fixture(scope="module")
def R():
    if not os.path.exists(oracle.REF_LIB):
        pytest.skip("oracle/_ref/libxw_ref.so not built (needs /root/reference)")
    L = C.CDLL(oracle.REF_LIB)
    if not hasattr(L, "ref_wire_encode_packet"):
        pytest.skip("oracle/_ref predates the wire entry points")
    for f in ("ref_wire_encode_packet", "ref_wire_decode_dump", "ref_wire_request", "ref_wire_reply_reset",
              "ref_wire_reply_take_actions", "ref_wire_reply_get_state", "ref_wire_reply_text"):
        getattr(L, f).restype = C.c_long
    L.ref_wire_request.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_long]
    L.ref_wire_reply_reset.argtypes = [C.c_int] * 3 + [C.c_size_t] * 3 + [C.c_double] * 3 + [C.c_void_p, C.c_long]
    L.ref_wire_reply_take_actions.argtypes = [C.c_float, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_long]
    L.ref_wire_reply_get_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long]
    L.ref_wire_reply_text.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_long]
    L.ref_wire_encode_packet.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long]
    L.ref_wire_decode_dump.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long]
    L.ref_wire_read_take_actions_reply.argtypes = [C.c_void_p, C.c_long, C.POINTER(C.c_float), C.POINTER(C.c_int64),
                                                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_long]
    return L


def ref_fields(d):
    arr, keep = wire._fields(d)  # same layout as RefField
    return C.cast(arr, C.c_void_p), keep


def ref_call(fn, *args):
    buf = (C.c_uint8 * (1 << 20))()
    n = fn(*args, buf, len(buf))
    assert 0 < n <= len(buf)
    return bytes(buf[:n])


def ref_dump(R, data):
    out = C.create_string_buffer(4 << 20)
    n = R.ref_wire_decode_dump(data, len(data), out, len(out))
    assert 0 < n <= len(out)
    return out.value.decode()


PACKETS = [
    {"action": 3},
    {"action": np.array([1, 2, 3], np.int32)},
    {"pred_sentence": "go to the apple please ."},
    {"screen": np.arange(3 * 7 * 5, dtype=np.uint8)},
    {"reward": np.array([0.25, -1.5, 3e-9], np.float32)},
    {"sentence": ""},
    {"action": 2, "pred_sentence": "hello world"},
    {"reward": np.array([-0.01], np.float32), "screen": np.full(84 * 84 * 3, 255, np.uint8), "sentence": "Well done !"},
    {},
]


def test_packets_against_the_reference_encoder(R):
    for d in PACKETS:
        mine = wire.encode_packet(d)
        p, keep = ref_fields(d)
        theirs = ref_call(R.ref_wire_encode_packet, p, len(d))
        if len(d) <= 1:
            assert mine == theirs, d.keys()          # byte for byte
        else:                                         # the reference walks an unordered_map: compare what decodes
            assert len(mine) == len(theirs)
            assert ref_dump(R, mine) == ref_dump(R, theirs), d.keys()
        back = wire.decode_packet(theirs)             # this decoder on the reference's bytes
        assert set(back) == set(d)
        for k, v in d.items():
            want = np.atleast_1d(np.asarray(v)) if not isinstance(v, str) else v
            assert (back[k] == want) if isinstance(v, str) else (np.asarray(back[k]) == want).all(), k
        assert wire.encode_packet(wire.decode_packet(mine)) == mine


def test_the_reference_serialization_test_vector(R):
    """tests/test_statepacket.cpp:77-104 (StatePacket, serialization): "screen" = pixels {1,2,3,4} + ids {10,11},
    "internal_state" = reals {1.5..6.5} + str "abc"; encode, decode, compare -- here through the C ABI on one side
    and the reference's own StatePacket on the other."""
    lib = _abi.load()
    px = (C.c_uint8 * 4)(1, 2, 3, 4)
    ids = (C.c_int32 * 2)(10, 11)
    re = (C.c_float * 6)(1.5, 2.5, 3.5, 4.5, 5.5, 6.5)
    f = (_abi.XwWireField * 2)()
    f[0].key, f[0].pixels, f[0].n_pixels, f[0].ids, f[0].n_ids = b"screen", px, 4, ids, 2
    f[1].key, f[1].reals, f[1].n_reals, f[1].str = b"internal_state", re, 6, b"abc"
    n = lib.xw_wire_encode_packet(f, 2, None, 0)
    buf = (C.c_uint8 * n)()
    assert lib.xw_wire_encode_packet(f, 2, buf, n) == n
    mine = bytes(buf)
    theirs = ref_call(R.ref_wire_encode_packet, C.cast(f, C.c_void_p), 2)
    assert len(mine) == len(theirs) and ref_dump(R, mine) == ref_dump(R, theirs)
    assert "screen|06" in ref_dump(R, mine) and "internal_state|09" in ref_dump(R, mine)  # flag bytes: pixels|id, reals|str
    out = (_abi.XwWireField * 8)()
    k, used = C.c_int32(), C.c_size_t()
    tb = (C.c_uint8 * len(theirs)).from_buffer_copy(theirs)
    assert lib.xw_wire_decode_packet(tb, len(theirs), out, 8, C.byref(k), C.byref(used)) == 0
    assert k.value == 2 and used.value == len(theirs)  # EXPECT_TRUE(buf.eof())
    got = {out[i].key: out[i] for i in range(2)}
    s, i = got[b"screen"], got[b"internal_state"]
    assert C.string_at(s.pixels, 4) == bytes(px) and s.n_ids == 2 and C.string_at(s.ids, 8) == bytes(ids) and not s.reals and s.str is None
    assert i.n_reals == 6 and C.string_at(i.reals, 24) == bytes(re) and i.str == b"abc" and not i.pixels and not i.ids


def test_requests_and_replies_byte_for_byte(R):
    # requests, as SimulatorServer sends them
    for act in ({"action": 1}, {"pred_sentence": "what"}):
        p, keep = ref_fields(act)
        theirs = ref_call(R.ref_wire_request, b"take_actions", p, len(act), 4, 1, 0.0)
        assert wire.compose_request("take_actions", act, act_rep=4, show_screen=True) == theirs
        req = wire.parse_request(theirs[8:])
        assert struct.unpack("<Q", theirs[:8])[0] == len(theirs) - 8
        assert (req["cmd"], req["act_rep"], req["show_screen"]) == ("take_actions", 4, True) and set(req["actions"]) == set(act)
    p, keep = ref_fields({"action": 2, "pred_sentence": "to the left"})
    req = wire.parse_request(ref_call(R.ref_wire_request, b"take_actions", p, 2, 1, 0, 0.0)[8:])
    assert int(req["actions"]["action"][0]) == 2 and req["actions"]["pred_sentence"] == "to the left"
    theirs = ref_call(R.ref_wire_request, b"get_state", None, 0, 0, 0, C.c_float(1.75))
    assert wire.compose_request("get_state", reward=1.75) == theirs and wire.parse_request(theirs[8:])["reward"] == 1.75
    for cmd in ("reset", "report_perf", "get_extra_info", "stop"):
        theirs = ref_call(R.ref_wire_request, cmd.encode(), None, 0, 0, 0, 0.0)
        assert wire.compose_request(cmd) == theirs and wire.parse_request(theirs[8:])["cmd"] == cmd
    # replies, as SimulatorClient sends them
    assert wire.reply_reset(4, 0, 1, 84, 84, 3, 8.0, 8.0, 0.0) == ref_call(R.ref_wire_reply_reset, 4, 0, 1, 84, 84, 3, 8.0, 8.0, 0.0)
    mine = wire.reply_take_actions(-0.01, 12345678901, 4, 0, True, "1")
    assert mine == ref_call(R.ref_wire_reply_take_actions, C.c_float(-0.01), 12345678901, 4, 0, 1, b"1")
    r, ns, go, lv, ok = C.c_float(), C.c_int64(), C.c_int(), C.c_int(), C.c_int()
    la = C.create_string_buffer(64)
    assert R.ref_wire_read_take_actions_reply(mine[8:], len(mine) - 8, C.byref(r), C.byref(ns), C.byref(go), C.byref(lv),
                                              C.byref(ok), la, 64) == 0
    assert (np.float32(r.value), ns.value, go.value, lv.value, ok.value, la.value) == (np.float32(-0.01), 12345678901, 4, 0, 1, b"1")
    st = {"screen": np.arange(200, dtype=np.uint8)}
    p, keep = ref_fields(st)
    assert wire.reply_get_state(st) == ref_call(R.ref_wire_reply_get_state, p, 1)
    assert wire.reply_text("get_extra_info", "17|task:,event:,height:8,width:8") == \
        ref_call(R.ref_wire_reply_text, b"get_extra_info", b"17|task:,event:,height:8,width:8")
    assert wire.reply_text("report_perf") == ref_call(R.ref_wire_reply_text, b"report_perf", None)


def test_empty_vectors_round_trip():
    """tests/test_binary_buffer.cpp:160-175: an empty vector travels as a count of 0 and reads back empty (the
    reference's StateBuffer::set_value CHECKs on an empty range, so only the codec itself is exercised here)."""
    for d in ({"reward": np.zeros(0, np.float32)}, {"action": np.zeros(0, np.int32)}, {"screen": np.zeros(0, np.uint8)}):
        data = wire.encode_packet(d)
        k = next(iter(d))
        assert len(data) == 8 + (8 + len(k) + 1) + 1 + 8
        back = wire.decode_packet(data)
        assert list(back) == [k] and len(back[k]) == 0 and back[k].dtype == d[k].dtype


def test_malformed_input_is_an_error_not_a_crash():
    good = wire.compose_request("take_actions", {"action": 1})[8:]
    for cut in range(len(good)):
        with pytest.raises(RuntimeError):
            wire.parse_request(good[:cut])
    with pytest.raises(RuntimeError):
        wire.decode_packet(struct.pack("<Q", 1 << 40))


class Trainer(object):
    """One SimulatorServer (simulator_interface.cpp:165-313): listens, then issues blocking remote calls."""

    def __init__(self):
        self.lsock = socket.socket()
        self.lsock.bind(("127.0.0.1", 0))
        self.lsock.listen(1)
        self.port = self.lsock.getsockname()[1]
        self.conn = None

    def accept(self):
        self.conn, _ = self.lsock.accept()

    def call(self, msg, expect=None):
        self.conn.sendall(msg)
        if expect is None:
            return None
        hdr = self._read(8)
        body = self._read(struct.unpack("<Q", hdr)[0])
        assert body[8:8 + len(expect)] == expect.encode()  # read_msg(reply); CHECK_EQ(reply, func_name)
        return body

    def _read(self, n):
        b = b""
        while len(b) < n:
            c = self.conn.recv(n - len(b))
            assert c
            b += c
        return b


def test_batch_client_serves_simple_game_over_tcp(R):
    """BASELINE config 1 game (host engine, runs without a GPU) behind 3 connections; the trainer's bytes come from
    the reference's composer.  Envs are driven one at a time AND together; results equal the oracle's SimpleGame."""
    n = 3
    sim = Simulator.create("simple_game", {"array_size": 8, "n_envs": n})
    trainers = [Trainer() for _ in range(n)]
    client = {}

    def run():
        client["c"] = wire.BatchClient(sim, [t.port for t in trainers])
        client["c"].serve()

    th = threading.Thread(target=run, daemon=True)
    th.start()
    for t in trainers:
        t.accept()
    L = oracle.lib()
    games = [oracle.XoSimpleGame() for _ in range(n)]
    for t, g in zip(trainers, games):  # reset, one env after the other
        body = t.call(ref_call(R.ref_wire_request, b"reset", None, 0, 0, 0, 0.0), "reset")
        num_actions, over, lives = struct.unpack("<iii", body[14:26])
        assert (num_actions, over, lives) == (2, 0, 1)
        L.xo_sg_reset(C.byref(g), 8)
    rng = np.random.RandomState(0)
    L.xo_sg_act.restype = C.c_float
    steps = [0] * n
    served = 0
    for it in range(12):
        who = [i for i in range(n) if rng.rand() < 0.7] or [0]
        acts = {i: int(rng.randint(0, 2)) for i in who}
        for i in who:  # requests go out together ...
            p, keep = ref_fields({"action": acts[i]})
            trainers[i].conn.sendall(ref_call(R.ref_wire_request, b"take_actions", p, 1, 1, 0, 0.0))
        for i in who:  # ... and every env answers its own
            hdr = trainers[i]._read(8)
            body = trainers[i]._read(struct.unpack("<Q", hdr)[0])
            r, ns, go, lv, ok = C.c_float(), C.c_int64(), C.c_int(), C.c_int(), C.c_int()
            la = C.create_string_buffer(64)
            assert R.ref_wire_read_take_actions_reply(body, len(body), C.byref(r), C.byref(ns), C.byref(go), C.byref(lv),
                                                      C.byref(ok), la, 64) == 0
            steps[i] += 1
            served += 1
            want = L.xo_sg_act(C.byref(games[i]), acts[i])
            assert np.float32(r.value) == np.float32(want) and ns.value == steps[i] and la.value == str(acts[i]).encode()
            assert (go.value != 0) == bool(L.xo_sg_game_over(C.byref(games[i]))) and lv.value == (0 if go.value else 1)
            if go.value:
                trainers[i].call(ref_call(R.ref_wire_request, b"reset", None, 0, 0, 0, 0.0), "reset")
                L.xo_sg_reset(C.byref(games[i]), 8)
                steps[i] = 0
    body = trainers[1].call(ref_call(R.ref_wire_request, b"get_state", None, 0, 0, 0, C.c_float(0.5)), "get_state")
    st = wire.decode_packet(body[8 + len("get_state") + 1:])
    assert list(st["screen"]) == list(games[1].state)[:8] and st["reward"][0] == np.float32(0.5)
    assert trainers[2].call(ref_call(R.ref_wire_request, b"report_perf", None, 0, 0, 0, 0.0), "report_perf") is not None
    for t in trainers:
        t.call(ref_call(R.ref_wire_request, b"stop", None, 0, 0, 0, 0.0))
    th.join(timeout=10)
    assert not th.is_alive() and client["c"].steps_served == served
    assert client["c"].batches <= served


def test_batch_client_survives_a_misbehaving_connection():
    """One connection sending garbage, an oversized frame, a take_actions without an action or an out-of-range action is
    closed; the other envs keep being served, and a message already buffered behind another one is not left waiting."""
    import socket
    n = 5
    sim = Simulator.create("simple_game", {"array_size": 8, "n_envs": n})
    lsocks = []
    for _ in range(n):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        s.listen(1)
        lsocks.append(s)
    box = {}

    def run():
        box["c"] = wire.BatchClient(sim, [s.getsockname()[1] for s in lsocks])
        box["c"].serve()

    th = threading.Thread(target=run, daemon=True)
    th.start()
    conns = [s.accept()[0] for s in lsocks]

    def read_msg(c):
        b = b""
        while len(b) < 8:
            d = c.recv(8 - len(b))
            assert d
            b += d
        size = struct.unpack("<Q", b)[0]
        body = b""
        while len(body) < size:
            body += c.recv(size - len(body))
        return body

    conns[0].sendall(wire.compose_request("reset"))
    assert read_msg(conns[0])[8:13] == b"reset"
    conns[1].sendall(struct.pack("<Q", 5) + b"\xff\xfe\xfd\xfc\xfb")                  # not a message
    conns[2].sendall(struct.pack("<Q", 1 << 40))                                       # absurd frame size
    conns[3].sendall(wire.compose_request("take_actions", {"pred_sentence": "hi"}))    # no "action"
    conns[4].sendall(wire.compose_request("take_actions", {"action": 7}))              # simple_game has 2 actions
    for c in conns[1:]:
        c.settimeout(5)
        assert c.recv(16) == b""                                                       # closed by the client
    # env 0 is still served; two messages sent back to back are both answered without further traffic
    conns[0].sendall(wire.compose_request("take_actions", {"action": 1}) + wire.compose_request("take_actions", {"action": 1}))
    assert read_msg(conns[0])[8:20] == b"take_actions" and read_msg(conns[0])[8:20] == b"take_actions"
    conns[0].sendall(wire.compose_request("take_actions", {"action": 1}) + wire.compose_request("stop"))
    assert read_msg(conns[0])[8:20] == b"take_actions"
    th.join(timeout=10)
    assert not th.is_alive()
    assert sorted(i for i, _ in box["c"].errors) == [1, 2, 3, 4] and box["c"].steps_served == 3
