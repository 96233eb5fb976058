"""MsgDrone wire format (server/DroneMsg.proto:2-11) written out by hand in quadsim_b200/stream.py against google.protobuf
itself: the schema is rebuilt from the .proto's field list with descriptor_pb2 (the reference's generated DroneMsg_pb2.py is
rejected by protobuf >= 4), a message is filled the way pub_server.send_state does (server/pub_server.py:26-36) and the two
serialisations must be byte-identical."""
import numpy as np
import pytest

from quadsim_b200 import stream


def _msg_class():
    pytest.importorskip("google.protobuf")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="DroneMsg_test.proto", syntax="proto3")
    m = fd.message_type.add(name="MsgDrone")
    F = descriptor_pb2.FieldDescriptorProto
    m.field.add(name="id", number=1, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    m.field.add(name="sim_step", number=2, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    for k, name in enumerate(("pos", "vel", "att_quat", "att_rate")):
        m.field.add(name=name, number=3 + k, type=F.TYPE_DOUBLE, label=F.LABEL_REPEATED)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("MsgDrone"))


def _reference_bytes(cls, drone_id, sim_step, state_now):
    msg = cls()  # pub_server.send_state, statement for statement
    msg.id = drone_id
    msg.sim_step = sim_step
    state = state_now.tolist()
    msg.pos[:] = state[0:3]
    msg.vel[:] = state[3:6]
    msg.att_quat[:] = state[6:10]
    msg.att_rate[:] = state[10:-2]
    return msg.SerializeToString()


@pytest.mark.parametrize("drone_id,sim_step", [(1, 1), (2, 599), (1, 0), (0, 7), (2, 300000), (1, -3)])
def test_encoding_is_byte_identical_to_protobuf(drone_id, sim_step):
    cls = _msg_class()
    rng = np.random.RandomState(sim_step & 0xFFFF)
    for _ in range(20):
        state = rng.normal(size=13) * rng.choice([1e-3, 1.0, 50.0])
        state[rng.randint(13)] = 0.0
        ours = stream.encode_drone_msg(drone_id, sim_step, state)
        assert ours == _reference_bytes(cls, drone_id, sim_step, state)
        back = cls.FromString(ours)
        assert back.id == drone_id and back.sim_step == sim_step and list(back.pos) == state[0:3].tolist()
        assert len(back.att_rate) == 1 and back.att_rate[0] == state[10]  # the reference's state[10:-2] slip, kept
        d = stream.decode_drone_msg(ours)
        assert d["id"] == drone_id and d["sim_step"] == sim_step and d["att_quat"] == state[6:10].tolist()
    full = stream.decode_drone_msg(stream.encode_drone_msg(1, 5, np.arange(13.0), literal_rates=False))
    assert full["att_rate"] == [10.0, 11.0, 12.0] and full["vel"] == [3.0, 4.0, 5.0]
