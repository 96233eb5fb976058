"""State streaming to the reference's Unity bridge.

The reference publishes one drone's state per step as a protobuf `MsgDrone` on a ZeroMQ PUB socket
(`server/pub_server.py:10-36`; schema `server/DroneMsg.proto:2-11`: `int32 id = 1; int32 sim_step = 2; repeated double
pos = 3, vel = 4, att_quat = 5, att_rate = 6`), chaser on tcp port 12345 and target on 12346.  Here a handful of envs are
picked out of the batch on the device (C ABI `qs_peek_state`, O(k) -- the batch is never exported) and encoded on the host.
The wire format is written out by hand (proto3: varint fields omitted when 0, repeated doubles packed), so neither the
reference's generated `DroneMsg_pb2.py` (rejected by protobuf >= 4) nor protobuf itself is needed to send; the bytes are
identical to `MsgDrone.SerializeToString()` (checked in tests against google.protobuf when it is installed)."""
import ctypes as C
import struct

import numpy as np
import torch

from . import _capi

CHASER_ID, TARGET_ID = 1, 2                    # run_PID_server.py / video_docking_env.py: srv(1) chaser, srv(2) target
CHASER_PORT, TARGET_PORT = 12345, 12346        # pub_server.py:21-24


def _varint(v):
    v &= (1 << 64) - 1  # negative int32 are sign-extended to 64 bits (10 bytes), as protobuf does
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _packed_doubles(field, values):
    values = [float(x) for x in values]
    if not values:
        return b""
    body = struct.pack("<%dd" % len(values), *values)
    return _varint((field << 3) | 2) + _varint(len(body)) + body


def encode_drone_msg(drone_id, sim_step, state, literal_rates=True):
    """`pub_server.send_state(sim_step, state_now)` (`pub_server.py:26-36`) -> the serialized MsgDrone bytes.
    `literal_rates`: the reference sends `state[10:-2]`, i.e. ONE of the three body rates of a 13-vector (a slicing slip:
    `[10:]` was meant); False sends all three."""
    state = np.asarray(state, dtype=np.float64).reshape(-1)
    out = b""
    if int(drone_id) != 0:
        out += _varint((1 << 3) | 0) + _varint(int(drone_id))
    if int(sim_step) != 0:
        out += _varint((2 << 3) | 0) + _varint(int(sim_step))
    out += _packed_doubles(3, state[0:3]) + _packed_doubles(4, state[3:6]) + _packed_doubles(5, state[6:10])
    out += _packed_doubles(6, state[10:-2] if literal_rates else state[10:])
    return out


def decode_drone_msg(data):
    """Inverse of encode_drone_msg (what the Unity client parses) -> dict(id, sim_step, pos, vel, att_quat, att_rate)."""
    msg = dict(id=0, sim_step=0, pos=[], vel=[], att_quat=[], att_rate=[])
    names = {1: "id", 2: "sim_step", 3: "pos", 4: "vel", 5: "att_quat", 6: "att_rate"}
    i = 0

    def varint():
        nonlocal i
        v, shift = 0, 0
        while True:
            b = data[i]
            i += 1
            v |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                return v

    while i < len(data):
        key = varint()
        field, wt = key >> 3, key & 7
        if wt == 0:
            v = varint()
            msg[names[field]] = v - (1 << 64) if v >> 63 else v
        elif wt == 2:
            n = varint()
            msg[names[field]] = list(struct.unpack("<%dd" % (n // 8), data[i:i + n]))
            i += n
        elif wt == 1:  # unpacked repeated double (proto2-style senders)
            msg[names[field]].append(struct.unpack("<d", data[i:i + 8])[0])
            i += 8
        else:
            raise ValueError("unexpected wire type %d" % wt)
    return msg


class StatePeek:
    """Device-side gather of a few envs' drone states (pinned host landing buffers, one small kernel + one copy)."""

    def __init__(self, env, env_indices=(0,)):
        self.env = env
        idx = [int(i) for i in np.atleast_1d(env_indices)]
        if any(i < 0 or i >= env.num_envs for i in idx):
            raise IndexError("env index out of range")
        self.k = len(idx)
        self.ids = torch.tensor(idx, dtype=torch.int64, device=env.device)
        self.docking = env.kind != _capi.HOVER
        self._dev = torch.empty((2, self.k, 13), dtype=torch.float64, device=env.device)
        self._t_dev = torch.empty((self.k,), dtype=torch.int32, device=env.device)
        self._host = torch.empty((2, self.k, 13), dtype=torch.float64, pin_memory=True)
        self._t_host = torch.empty((self.k,), dtype=torch.int32, pin_memory=True)

    def read(self):
        """-> (chaser[k,13], target[k,13] or None, t[k]) numpy, world coordinates, float64 whatever the env dtype."""
        env = self.env
        env._check_open()
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(env.device):
            _capi.check(env.lib.qs_peek_state(env._h, p(self.ids), self.k, p(self._dev[0]), p(self._dev[1]) if self.docking else None,
                                              p(self._t_dev), env._stream()))
            self._host.copy_(self._dev, non_blocking=True)
            self._t_host.copy_(self._t_dev, non_blocking=True)
            torch.cuda.current_stream(env.device).synchronize()
        return self._host[0].numpy().copy(), (self._host[1].numpy().copy() if self.docking else None), self._t_host.numpy().copy()


class StatePublisher:
    """`pub_server` x2 for one env of the batch: call `send()` after `env.step` to publish the chaser (id 1) and, for
    docking, the target (id 2) exactly as `run_PID_server.py` / `video_docking_env.py` do.  `endpoints` default to the
    reference's `tcp://*:12345` / `tcp://*:12346`; `sim_step` defaults to the env's own step counter."""

    def __init__(self, env, env_index=0, endpoints=None, literal_rates=True, context=None):
        try:
            import zmq
        except ImportError as exc:  # the transport is the reference's; without it there is nothing to publish on
            raise _capi.QuadSimError("StatePublisher needs pyzmq (the reference's transport, server/pub_server.py:1)") from exc
        self.peek = StatePeek(env, (env_index,))
        self.literal_rates = literal_rates
        self._own = context is None
        self.context = zmq.Context() if context is None else context
        eps = endpoints or ("tcp://*:%d" % CHASER_PORT, "tcp://*:%d" % TARGET_PORT)
        self.sockets = []
        for ep in eps[:2 if self.peek.docking else 1]:
            s = self.context.socket(zmq.PUB)
            s.bind(ep)
            self.sockets.append(s)

    def send(self, sim_step=None):
        chaser, target, t = self.peek.read()
        step = int(t[0]) if sim_step is None else int(sim_step)
        msgs = [encode_drone_msg(CHASER_ID, step, chaser[0], self.literal_rates)]
        if target is not None:
            msgs.append(encode_drone_msg(TARGET_ID, step, target[0], self.literal_rates))
        for s, m in zip(self.sockets, msgs):
            s.send(m)
        return msgs

    def close(self):
        for s in self.sockets:
            s.close(linger=0)
        self.sockets = []
        if self._own:
            self.context.term()
