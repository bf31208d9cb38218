"""ROS1 wire formats of the messages either side of the hot path - Python mirror of
include/cngp_wire.hpp, byte-compatible with it (tests/test_wire.py).

    core_nav/GP_Input      core_navigation/msg/GP_Input.msg:1-3     time_array / slip_array into gp_slip_node.callback
    core_nav/GP_Output     core_navigation/msg/GP_Output.msg:1-3    mean / sigma out of it (gp_slip_node.py:57-63)
    core_nav/SetStopping   core_navigation/srv/SetStopping.srv:1-7  the look-ahead context (gp_predictor.cpp:26-46)
    std_msgs/Float64       stop_cmd (gp_predictor.cpp:118)

rospy and genpy are not in this image; the layout is restated from the ROS1 serialisation rules: little-endian, uint32
length prefix for strings and variable-length arrays, none for fixed-length arrays, Header = seq, (secs, nsecs),
frame_id, bool = one byte.  `md5sum` rebuilds the ROS "md5 text" sum of a definition."""
from __future__ import annotations

import hashlib
import math
import struct
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

MD5 = {
    "std_msgs/Header": "2176decaecbce78abc3b96ef049fabed",
    "geometry_msgs/Point": "4a842b65f413084dc2b10fb484ea7f17",
    "std_msgs/Float64": "fdb28210bfa9d7c91146260178d9a584",
    "core_nav/GP_Input": "9753e28f26b0947dec1baef0e82339bc",
    "core_nav/GP_Output": "aa85e91d502deb241dc28762eb372b44",
    "core_nav/SetStopping": "24fce43738a51f1ac343c3c21c375939",
}
DEFINITIONS = {
    "std_msgs/Header": "uint32 seq\ntime stamp\nstring frame_id",
    "geometry_msgs/Point": "float64 x\nfloat64 y\nfloat64 z",
    "std_msgs/Float64": "float64 data",
    "core_nav/GP_Input": "Header header\nfloat64[] time_array\nfloat64[] slip_array",
    "core_nav/GP_Output": "Header header\nfloat64[] mean\nfloat64[] sigma",
    "core_nav/SetStopping": ("bool stopping", "float64[225] PvecData\nfloat64[225] QvecData\nfloat64[225] STMvecData\n"
                             "float64[60] HvecData\ngeometry_msgs/Point PosData"),
}
_BUILTIN = {"bool", "int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64", "float32", "float64",
            "string", "time", "duration"}
SET_STOPPING_RESPONSE_BYTES = (3 * 225 + 60 + 3) * 8


def _md5_text(text: str) -> str:
    out = []
    for line in text.split("\n"):
        typ, name = line.split()
        base = typ.split("[")[0]
        if base not in _BUILTIN:
            full = "std_msgs/Header" if base == "Header" else base
            typ = md5sum(full)           # embedded messages are replaced by their own sum (arrays lose the brackets)
        out.append(f"{typ} {name}")
    return "\n".join(out)


def md5sum(name: str) -> str:
    """ROS md5 of a message (md5 of its md5 text) or service (md5 of request text + response text)."""
    d = DEFINITIONS[name]
    text = _md5_text(d) if isinstance(d, str) else _md5_text(d[0]) + _md5_text(d[1])
    return hashlib.md5(text.encode()).hexdigest()


@dataclass
class Header:
    """std_msgs/Header.  `stamp` (seconds, float64) is the convenient view; `secs` / `nsecs` hold the integer pair a
    message arrived with.  A float64 at epoch scale resolves ~240 ns, so re-deriving the pair from the float can change
    nsecs: serialisation re-uses the integer pair whenever `stamp` still equals the float it was decoded to, which
    makes deserialise -> serialise byte-exact (ExactTime matching, republishing)."""
    seq: int = 0
    stamp: float = 0.0
    frame_id: str = ""
    secs: Optional[int] = field(default=None, compare=False)
    nsecs: Optional[int] = field(default=None, compare=False)


@dataclass
class GPInput:
    header: Header = field(default_factory=Header)
    time_array: np.ndarray = field(default_factory=lambda: np.zeros(0))
    slip_array: np.ndarray = field(default_factory=lambda: np.zeros(0))


@dataclass
class GPOutput:
    header: Header = field(default_factory=Header)
    mean: np.ndarray = field(default_factory=lambda: np.zeros(0))
    sigma: np.ndarray = field(default_factory=lambda: np.zeros(0))


@dataclass
class SetStoppingResponse:
    PvecData: np.ndarray = field(default_factory=lambda: np.zeros(225))
    QvecData: np.ndarray = field(default_factory=lambda: np.zeros(225))
    STMvecData: np.ndarray = field(default_factory=lambda: np.zeros(225))
    HvecData: np.ndarray = field(default_factory=lambda: np.zeros(60))
    PosData: np.ndarray = field(default_factory=lambda: np.zeros(3))


def stamp_to_ros(t: float):
    """rospy.Time.from_sec / ros::Time(double): secs = floor(t), nsecs = round((t - secs) 1e9), normalised."""
    fl = math.floor(t)
    s, ns = int(fl), int(round((t - fl) * 1e9))
    s, ns = s + ns // 1_000_000_000, ns % 1_000_000_000
    if not 0 <= s <= 0xFFFFFFFF:
        raise ValueError("stamp outside the ROS time range")
    return s, ns


def _put_header(h: Header) -> bytes:
    if h.secs is not None and h.nsecs is not None and h.stamp == h.secs + 1e-9 * h.nsecs:
        s, ns = int(h.secs), int(h.nsecs)          # untouched since it was decoded: keep the exact pair
    else:
        s, ns = stamp_to_ros(h.stamp)
    fid = h.frame_id.encode()
    return struct.pack("<IIII", h.seq, s, ns, len(fid)) + fid


def _get_header(b: memoryview, at: int):
    seq, s, ns, k = struct.unpack_from("<IIII", b, at)
    at += 16
    if at + k > len(b):
        raise ValueError("truncated message")
    return Header(seq, s + 1e-9 * ns, bytes(b[at:at + k]).decode(), secs=s, nsecs=ns), at + k


def _put_vec(v) -> bytes:
    a = np.ascontiguousarray(v, dtype="<f8").reshape(-1)
    return struct.pack("<I", a.size) + a.tobytes()


def _get_vec(b: memoryview, at: int):
    (k,) = struct.unpack_from("<I", b, at)
    at += 4
    if at + 8 * k > len(b):
        raise ValueError("truncated message")
    return np.frombuffer(b, dtype="<f8", count=k, offset=at).copy(), at + 8 * k


def _two_vectors(h, a, b) -> bytes:
    return _put_header(h) + _put_vec(a) + _put_vec(b)


def serialize(m) -> bytes:
    if isinstance(m, GPInput):
        return _two_vectors(m.header, m.time_array, m.slip_array)
    if isinstance(m, GPOutput):
        return _two_vectors(m.header, m.mean, m.sigma)
    if isinstance(m, SetStoppingResponse):
        parts = [(m.PvecData, 225), (m.QvecData, 225), (m.STMvecData, 225), (m.HvecData, 60), (m.PosData, 3)]
        out = b""
        for v, n in parts:
            a = np.ascontiguousarray(v, dtype="<f8").reshape(-1)
            if a.size != n:
                raise ValueError(f"fixed-length field has {a.size} entries, expected {n}")
            out += a.tobytes()
        return out
    if isinstance(m, bool):                         # SetStopping request
        return b"\x01" if m else b"\x00"
    if isinstance(m, float):                        # std_msgs/Float64
        return struct.pack("<d", m)
    raise TypeError(type(m))


def _deserialize_two(cls, data):
    b = memoryview(data)
    h, at = _get_header(b, 0)
    v0, at = _get_vec(b, at)
    v1, at = _get_vec(b, at)
    if at != len(b):
        raise ValueError("trailing bytes")
    return cls(h, v0, v1)


def deserialize_gp_input(data) -> GPInput:
    return _deserialize_two(GPInput, data)


def deserialize_gp_output(data) -> GPOutput:
    return _deserialize_two(GPOutput, data)


def deserialize_set_stopping_response(data) -> SetStoppingResponse:
    if len(data) != SET_STOPPING_RESPONSE_BYTES:
        raise ValueError(f"SetStopping response is {SET_STOPPING_RESPONSE_BYTES} bytes, got {len(data)}")
    a = np.frombuffer(data, dtype="<f8").copy()
    return SetStoppingResponse(a[0:225], a[225:450], a[450:675], a[675:735], a[735:738])


def deserialize_set_stopping_request(data) -> bool:
    if len(data) != 1:
        raise ValueError("SetStopping request is one byte")
    return data[0] != 0


def deserialize_float64(data) -> float:
    if len(data) != 8:
        raise ValueError("std_msgs/Float64 is eight bytes")
    return struct.unpack("<d", data)[0]


def frame(body: bytes) -> bytes:
    """TCPROS framing: uint32 byte count, then the serialised message."""
    return struct.pack("<I", len(body)) + body


def unframe(buf: bytes):
    """(body, bytes consumed) of the first whole frame in buf, or (None, 0)."""
    if len(buf) < 4:
        return None, 0
    (k,) = struct.unpack_from("<I", buf, 0)
    if len(buf) < 4 + k:
        return None, 0
    return bytes(buf[4:4 + k]), 4 + k


def connection_header(fields: dict) -> bytes:
    """TCPROS connection header: framed list of framed 'key=value' strings (callerid, topic, type, md5sum ...)."""
    body = b"".join(frame(f"{k}={v}".encode()) for k, v in fields.items())
    return frame(body)


def parse_connection_header(buf: bytes) -> dict:
    body, _ = unframe(buf)
    if body is None:
        raise ValueError("incomplete connection header")
    out, at = {}, 0
    while at < len(body):
        item, used = unframe(body[at:])
        if item is None:
            raise ValueError("malformed connection header")
        k, _, v = item.decode().partition("=")
        out[k] = v
        at += used
    return out
