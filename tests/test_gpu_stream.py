"""GPU check of the Unity-bridge streaming path: qs_peek_state picks envs out of the batch bit-identically to
the full export, and StatePublisher's ZeroMQ messages decode to the chaser / target states of the chosen env."""
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import quadsim_b200 as qs
    from quadsim_b200 import stream


@pytest.mark.parametrize("env_id,dtype", [("docking-v2", torch.float32), ("docking-v0", torch.float64), ("hovering-v0", torch.float32)])
def test_peek_equals_export(env_id, dtype):
    n = 1000
    env = qs.make_vec(env_id, n, dtype=dtype, seed=9, randomize="episode" if env_id != "hovering-v0" else None)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(37):
        env.step(torch.rand((n, 4), device="cuda", generator=g, dtype=dtype) * (1 if env_id == "hovering-v0" else 2) - (0 if env_id == "hovering-v0" else 1))
    idx = [0, 1, 127, 128, 999, 500]
    chaser, target, t = stream.StatePeek(env, idx).read()
    st = env.get_state()
    assert np.array_equal(chaser, st["chaser"].cpu().numpy()[idx]) and np.array_equal(t, st["t"].cpu().numpy()[idx])
    if env_id != "hovering-v0":
        assert np.array_equal(target, st["target"].cpu().numpy()[idx])
    else:
        assert target is None
    with pytest.raises(IndexError):
        stream.StatePeek(env, [n])
    env.close()


def test_publisher_round_trip_over_zmq():
    zmq = pytest.importorskip("zmq")
    env = qs.make_vec("docking-v0", 256, dtype=torch.float64, seed=1, randomize="episode")
    env.reset()
    ctx = zmq.Context()
    eps = ("inproc://qs-chaser", "inproc://qs-target")
    pub = stream.StatePublisher(env, env_index=200, endpoints=eps, context=ctx)
    subs = []
    for ep in eps:
        s = ctx.socket(zmq.SUB)
        s.connect(ep)
        s.setsockopt(zmq.SUBSCRIBE, b"")
        subs.append(s)
    time.sleep(0.2)  # PUB/SUB slow joiner
    act = torch.zeros((256, 4), dtype=torch.float64, device="cuda")
    for step in range(1, 4):
        env.step(act)
        sent = pub.send()
        st = env.get_state()
        for s, key, did, raw in zip(subs, ("chaser", "target"), (1, 2), sent):
            assert s.poll(2000), "no message"
            data = s.recv()
            assert data == raw
            m = stream.decode_drone_msg(data)
            x = st[key][200].cpu().numpy()
            assert m["id"] == did and m["sim_step"] == step
            assert m["pos"] == x[0:3].tolist() and m["vel"] == x[3:6].tolist() and m["att_quat"] == x[6:10].tolist()
            assert m["att_rate"] == x[10:11].tolist()
    pub.close()
    for s in subs:
        s.close(linger=0)
    ctx.term()
    env.close()
