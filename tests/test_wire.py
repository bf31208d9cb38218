"""ROS1 wire formats: corenav_gp_b200/wire.py and include/cngp_wire.hpp.

Known answers: the md5 method is pinned on the published sums of std_msgs/Header, geometry_msgs/Point and
std_msgs/Float64; the byte layout on hand-assembled messages (ROS1 serialisation rules); the C++ and Python sides must
produce identical bytes for the same message."""
import os
import struct
import subprocess

import numpy as np
import pytest

from corenav_gp_b200 import wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "host", "_build")


@pytest.fixture(scope="module")
def cli():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "wire_cli")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "wire_cli.cpp")])
    return exe


def through_cpp(cli, kind, data, tmp_path, expect_rc=0):
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    src.write_bytes(data)
    r = subprocess.run([cli, kind, str(src), str(dst)], capture_output=True, text=True, timeout=60)
    assert r.returncode == expect_rc, r.stderr
    return (dst.read_bytes() if expect_rc == 0 else None), r.stdout


def test_md5_method_reproduces_published_sums():
    assert wire.md5sum("std_msgs/Header") == "2176decaecbce78abc3b96ef049fabed"
    assert wire.md5sum("geometry_msgs/Point") == "4a842b65f413084dc2b10fb484ea7f17"
    assert wire.md5sum("std_msgs/Float64") == "fdb28210bfa9d7c91146260178d9a584"
    for name in ("core_nav/GP_Input", "core_nav/GP_Output", "core_nav/SetStopping"):
        assert wire.md5sum(name) == wire.MD5[name]
    hpp = open(os.path.join(ROOT, "include", "cngp_wire.hpp")).read()
    for name in wire.MD5:
        assert wire.MD5[name] in hpp


def test_gp_input_hand_assembled_bytes():
    m = wire.GPInput(wire.Header(7, 12.5, "odom"), np.array([21.0, 22.0]), np.array([0.25]))
    want = (struct.pack("<I", 7) + struct.pack("<II", 12, 500000000) + struct.pack("<I", 4) + b"odom"
            + struct.pack("<I", 2) + struct.pack("<dd", 21.0, 22.0) + struct.pack("<I", 1) + struct.pack("<d", 0.25))
    assert wire.serialize(m) == want
    back = wire.deserialize_gp_input(want)
    assert back.header == m.header and np.array_equal(back.time_array, m.time_array)
    assert np.array_equal(back.slip_array, m.slip_array)


def test_stamp_rounding_and_carry():
    assert wire.stamp_to_ros(1.9999999996) == (2, 0)          # nsecs rounds up to 1e9 and carries
    assert wire.stamp_to_ros(0.0) == (0, 0)
    assert wire.stamp_to_ros(1602931200.123456789)[0] == 1602931200
    with pytest.raises(ValueError):
        wire.stamp_to_ros(-1.0)


def test_empty_and_ragged_arrays_round_trip():
    for n, m in ((0, 0), (0, 3), (149, 149), (1, 0)):
        msg = wire.GPOutput(wire.Header(1, 3.0, ""), np.arange(n, dtype=float), np.arange(m, dtype=float) * 0.5)
        back = wire.deserialize_gp_output(wire.serialize(msg))
        assert back.mean.size == n and back.sigma.size == m
        assert np.array_equal(back.mean, msg.mean) and np.array_equal(back.sigma, msg.sigma)


def test_truncated_and_trailing_bytes_are_rejected():
    data = wire.serialize(wire.GPInput(wire.Header(1, 1.0, "a"), np.ones(4), np.ones(4)))
    with pytest.raises((ValueError, struct.error)):
        wire.deserialize_gp_input(data[:-3])
    with pytest.raises(ValueError):
        wire.deserialize_gp_input(data + b"\x00")
    with pytest.raises(ValueError):
        wire.deserialize_set_stopping_response(b"\x00" * 100)


def test_set_stopping_layout():
    rng = np.random.default_rng(0)
    r = wire.SetStoppingResponse(rng.standard_normal(225), rng.standard_normal(225), rng.standard_normal(225),
                                 rng.standard_normal(60), rng.standard_normal(3))
    data = wire.serialize(r)
    assert len(data) == wire.SET_STOPPING_RESPONSE_BYTES == 5904
    flat = np.frombuffer(data, dtype="<f8")
    assert np.array_equal(flat[675:735], r.HvecData) and np.array_equal(flat[735:], r.PosData)
    back = wire.deserialize_set_stopping_response(data)
    assert np.array_equal(back.STMvecData, r.STMvecData)
    assert wire.serialize(True) == b"\x01" and wire.deserialize_set_stopping_request(b"\x00") is False
    with pytest.raises(ValueError):
        wire.serialize(wire.SetStoppingResponse(PvecData=np.zeros(224)))


def test_framing_and_connection_header():
    body = wire.serialize(2.5)
    assert wire.frame(body) == b"\x08\x00\x00\x00" + struct.pack("<d", 2.5)
    assert wire.unframe(wire.frame(body)[:7]) == (None, 0)
    assert wire.unframe(wire.frame(body) + b"xx") == (body, 12)
    hdr = {"callerid": "/gp_predictor_node", "topic": "/core_nav/core_nav/gp_result", "type": "core_nav/GP_Output",
           "md5sum": wire.MD5["core_nav/GP_Output"]}
    assert wire.parse_connection_header(wire.connection_header(hdr)) == hdr


def test_cpp_and_python_agree_byte_for_byte(cli, tmp_path):
    rng = np.random.default_rng(1)
    gi = wire.serialize(wire.GPInput(wire.Header(42, 1234.000000001, "base_link"), 20.0 + np.arange(149), rng.standard_normal(149)))
    out, summary = through_cpp(cli, "gp_input", gi, tmp_path)
    assert out == gi and '"n": 149' in summary and '"frame_id": "base_link"' in summary
    go = wire.serialize(wire.GPOutput(wire.Header(3, 0.1, ""), rng.standard_normal(599), np.abs(rng.standard_normal(599))))
    assert through_cpp(cli, "gp_output", go, tmp_path)[0] == go
    sr = wire.serialize(wire.SetStoppingResponse(rng.standard_normal(225), rng.standard_normal(225), rng.standard_normal(225),
                                                 rng.standard_normal(60), np.array([0.69, -1.39, 334.9])))
    out, summary = through_cpp(cli, "set_stopping_response", sr, tmp_path)
    assert out == sr and '"z": 334.8999999' in summary
    assert through_cpp(cli, "set_stopping_request", b"\x01", tmp_path)[0] == b"\x01"
    assert through_cpp(cli, "float64", wire.serialize(0.5), tmp_path)[0] == wire.serialize(0.5)
    stream = wire.frame(go) + wire.frame(wire.serialize(wire.GPOutput()))
    assert through_cpp(cli, "framed_gp_output", stream, tmp_path)[0] == stream
    through_cpp(cli, "gp_output", go[:-1], tmp_path, expect_rc=4)        # truncated: std::out_of_range, reported


def test_wire_to_callback_to_wire(monkeypatch):
    """gp_slip_node.callback_bytes: GP_Input bytes -> callback -> GP_Output bytes.  The GP arithmetic is stubbed here (no
    GPU on this side; tests/test_gpu_fit_callback.py runs the real callback): what is checked is the byte plumbing."""
    from corenav_gp_b200 import gp_slip_node as node
    n = 40
    seen = {}

    def fake_callback(data, theta=None, kernel=node.KERNEL):
        seen["n"] = len(data.time_array)
        seen["t0"] = float(data.time_array[0])
        return node.GP_Output(mean=np.arange(599.0), sigma=np.full(599, 0.25))

    monkeypatch.setattr(node, "callback", fake_callback)
    gi = wire.GPInput(wire.Header(11, 5.5, "odom"), 20.0 + np.arange(n), 0.05 * np.sin(np.arange(n) / 5.0))
    reply = node.callback_bytes(wire.serialize(gi))
    out = wire.deserialize_gp_output(reply)
    assert seen == {"n": n, "t0": 20.0}
    assert out.header == gi.header and out.mean[598] == 598.0 and out.sigma.sum() == 599 * 0.25
    framed = node.callback_bytes(wire.frame(wire.serialize(gi)), framed=True)
    assert wire.unframe(framed) == (reply, len(reply) + 4)
    with pytest.raises(ValueError):
        node.callback_bytes(wire.frame(wire.serialize(gi))[:-1], framed=True)


def test_epoch_scale_stamp_round_trips_byte_exact():
    """A float64 at epoch scale resolves ~240 ns: deserialise -> serialise must keep the integer (secs, nsecs) pair."""
    import struct
    body = struct.pack("<IIII", 7, 1760716800, 123456789, 4) + b"base" + struct.pack("<I", 0) + struct.pack("<I", 0)
    m = wire.deserialize_gp_input(body)
    assert (m.header.secs, m.header.nsecs) == (1760716800, 123456789)
    assert wire.serialize(m) == body                                      # exact pair kept
    assert wire.stamp_to_ros(m.header.stamp)[1] != 123456789              # ... which the float alone cannot give back
    m.header.stamp += 1.0                                                 # an edited stamp is re-derived from the float
    assert struct.unpack_from("<II", wire.serialize(m), 4)[0] == 1760716801


def test_output_buffers_are_never_silently_replaced():
    from corenav_gp_b200.api import CngpError, _Arg
    import numpy as np
    ok = np.empty((3, 4))
    assert _Arg(ok, np.float64, False, output=True).ptr == ok.ctypes.data
    for bad in (np.empty((3, 4), dtype=np.float32), np.empty((4, 6))[:, ::2], [0.0] * 4):
        with pytest.raises(CngpError):
            _Arg(bad, np.float64, False, output=True)
    assert _Arg(np.empty((4, 6))[:, ::2], np.float64, False).keep.flags.c_contiguous       # inputs are coerced
