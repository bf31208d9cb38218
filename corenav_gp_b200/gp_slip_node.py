"""ROS-free mirror of the reference's GP node (core_navigation/script/gp_slip_node.py:12-83).

Same module-level surface as the reference script - `pub`, `callback(data)`, `gaussian_process()` - with rospy's
publisher / subscriber replaced by plain callables, and GPy replaced by one call into libcngp
(cngp_gp_slip_batch: train split, L-BFGS-B fit from all-ones, grid, predict, mean[n:], 2 sqrt(var[n:]) - all arithmetic
on the GPU).  A ROS deployment keeps its own thin script: subscribe `/core_nav/core_nav/gp_input`, call `callback`,
publish the returned message on `/core_nav/core_nav/gp_result` (INTEGRATION.md shows the stub).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from .api import GpContext

KERNEL = "rbf*brownian"      # gp_slip_node.py:31  kernel = GPy.kern.RBF(1) * GPy.kern.Brownian(1)
HORIZON = 600                # gp_slip_node.py:45


@dataclass
class Header:
    seq: int = 0
    stamp: float = 0.0
    frame_id: str = ""


@dataclass
class GP_Input:              # core_navigation/msg/GP_Input.msg:1-3
    header: Header = field(default_factory=Header)
    time_array: Sequence[float] = ()
    slip_array: Sequence[float] = ()


@dataclass
class GP_Output:             # core_navigation/msg/GP_Output.msg:1-3
    header: Header = field(default_factory=Header)
    mean: Sequence[float] = ()
    sigma: Sequence[float] = ()


class Publisher:
    """Stand-in for rospy.Publisher('/core_nav/core_nav/gp_result', GP_Output, queue_size=1): keeps the last message and
    forwards to an optional sink."""

    def __init__(self, sink: Optional[Callable[[GP_Output], None]] = None):
        self.sink = sink
        self.last: Optional[GP_Output] = None

    def publish(self, msg: GP_Output) -> None:
        self.last = msg
        if self.sink is not None:
            self.sink(msg)


pub = Publisher()
_ctx: Optional[GpContext] = None


def context(device: int = 0) -> GpContext:
    global _ctx
    if _ctx is None:
        _ctx = GpContext(device=device)
    return _ctx


class NotPositiveDefiniteError(ArithmeticError):
    """Ky of a window is not positive definite (NaN slip samples, a bad user-supplied theta).  In the reference GPy
    raises LinAlgError inside the callback and nothing is published (gp_slip_node.py:35-36); the C++ twin
    gp_slip_predict throws on the same condition."""


def _raise_on_failed(status) -> None:
    bad = np.flatnonzero(np.asarray(status) < 0)
    if bad.size:
        raise NotPositiveDefiniteError(f"GP window(s) {bad.tolist()[:8]} not positive definite (status "
                                       f"{np.asarray(status)[bad].tolist()[:8]}): nothing published")


def callback(data, theta=None, kernel: str = KERNEL) -> GP_Output:
    """gp_slip_node.callback: one GP_Input window in, one GP_Output message out (also handed to `pub`).

    theta=None fits the hyper-parameters first, as the reference does (m.optimize()); a fixed theta (kernel
    hyper-parameters then noise variance) skips the fit."""
    X = np.asarray(data.time_array, dtype=np.float64).reshape(1, -1)
    Y = np.asarray(data.slip_array, dtype=np.float64).reshape(1, -1)
    mean, sigma, status = context().gp_slip(kernel, X, Y, theta=theta, horizon=HORIZON)
    _raise_on_failed(status)
    msg_out = GP_Output()
    msg_out.mean = mean[0]
    msg_out.sigma = sigma[0]
    pub.publish(msg_out)
    return msg_out


def callback_batch(windows: Sequence, theta=None, kernel: str = KERNEL) -> List[GP_Output]:
    """Many windows of equal length in one GPU pass (Monte-Carlo use, BASELINE.json configs[3])."""
    X = np.stack([np.asarray(w.time_array, dtype=np.float64) for w in windows])
    Y = np.stack([np.asarray(w.slip_array, dtype=np.float64) for w in windows])
    mean, sigma, status = context().gp_slip(kernel, X, Y, theta=theta, horizon=HORIZON)
    _raise_on_failed(status)
    return [GP_Output(mean=mean[b], sigma=sigma[b]) for b in range(len(windows))]


def callback_bytes(serialized: bytes, theta=None, kernel: str = KERNEL, framed: bool = False) -> bytes:
    """The node as a byte pipe for a TCPROS / rosbag driver: a serialised core_nav/GP_Input in, the
    serialised core_nav/GP_Output of `callback` out (header copied, as a republisher would).  framed=True: both carry
    the uint32 TCPROS length prefix."""
    from . import wire
    if framed:
        body, used = wire.unframe(serialized)
        if body is None or used != len(serialized):
            raise ValueError("callback_bytes: expected exactly one whole TCPROS frame")
        serialized = body
    msg = wire.deserialize_gp_input(serialized)
    out = callback(GP_Input(time_array=msg.time_array, slip_array=msg.slip_array), theta=theta, kernel=kernel)
    reply = wire.serialize(wire.GPOutput(wire.Header(msg.header.seq, msg.header.stamp, msg.header.frame_id, msg.header.secs, msg.header.nsecs),
                                         np.asarray(out.mean), np.asarray(out.sigma)))
    return wire.frame(reply) if framed else reply


def gaussian_process(subscribe: Callable[[Callable], None]) -> None:
    """gp_slip_node.gaussian_process: attach `callback` to a message source (rospy.Subscriber in a ROS graph)."""
    subscribe(callback)
