"""CPU: oracle/slip_oracle.c (slip extraction + GP_Input recorder, CoreNav.cpp:244-329) against an independent numpy /
pure-Python transcription of the same source lines, plus the recorder's edge cases."""
import math

import numpy as np

from corenav_gp_b200 import synthetic as syn
from oracle import slip_oracle as so


def transcription(joint, att, vel, cmd, stop_cmd, max_windows, cap):
    """Line-by-line Python reading of CoreNav.cpp:176-183, 190, 244-329 for one drive."""
    T = len(cmd)
    count = start = stop = 0.0
    first, gp_flag, new_stop, cmd_stop = True, False, False, 0.0
    slips, windows, cur = [], [], []
    for k in range(T):
        if stop_cmd is not None and not math.isnan(stop_cmd[k]):
            cmd_stop, new_stop = stop_cmd[k], True
        count += 1
        vFL, vFR, vBL, vBR = -joint[k][0] * 0.11, joint[k][1] * 0.11, -joint[k][2] * 0.11, joint[k][3] * 0.11
        rear = (vBL + vBR) / 2.0
        C = syn.eul_to_dcm(*att[k])
        vlin = float((C @ vel[k])[0])
        with np.errstate(divide="ignore", invalid="ignore"):
            q = [np.float64(vFR - vlin) / np.float64(vFR), np.float64(vBR - vlin) / np.float64(vBR),
                 np.float64(vFL - vlin) / np.float64(vFL), np.float64(vBL - vlin) / np.float64(vBL)]
        mx = lambda a, b: b if a < b else a
        slip = float(mx(mx(q[0], q[1]), mx(q[2], q[3])))
        if abs(rear) < 0.001:
            slip = 0.0
        slip = max(-1.0, slip) if slip < -1.0 else slip
        slip = 1.0 if slip > 1.0 else slip
        slips.append(slip)
        if slip != 0.0 and slip != -1.0 and slip != 1.0 and abs(cmd[k]) > 0.2:
            if first:
                start, stop, first = count + 10, count + 10 + 150, False
            if start < count < stop and not gp_flag:
                cur.append((count, slip))
            if count >= stop:
                if not gp_flag:
                    gp_flag = True
                    windows.append((cur, len(cur) >= 15, k))
                    cur = []
                if new_stop:
                    new_stop = False
                    start = stop + math.ceil(cmd_stop) * 10 + 10 + 50
                    stop = start + 150
                    gp_flag = False
            if not first and count / 10 - stop / 10 > 10:
                cur, first, gp_flag = [], True, False
    return slips, windows


def check(d, max_windows=3, cap=149):
    out = so.slip_record(d["joint"], d["att"], d["vel"], d["cmd"], d["stop_cmd"], max_windows, cap)
    B = d["cmd"].shape[0]
    for b in range(B):
        slips, windows = transcription(d["joint"][b], d["att"][b], d["vel"][b], d["cmd"][b],
                                       None if d["stop_cmd"] is None else d["stop_cmd"][b], max_windows, cap)
        np.testing.assert_allclose(out["slip"][b], slips, rtol=1e-12, atol=1e-14, equal_nan=True)   # slip is a difference of O(1) terms
        assert out["n_windows"][b] == len(windows)
        for w, (cur, pub, k) in enumerate(windows[:max_windows]):
            assert out["n_samples"][b, w] == len(cur) and out["published"][b, w] == int(pub)
            assert out["stop_update"][b, w] == k
            n = min(len(cur), cap)
            assert np.array_equal(out["time_array"][b, w, :n], [c for c, _ in cur[:n]])
            np.testing.assert_allclose(out["slip_array"][b, w, :n], [s for _, s in cur[:n]], rtol=1e-12, atol=1e-14)
    return out


def test_synthetic_drives_match_transcription():
    d = syn.drives(0, 24, T=420)
    out = check(d)
    assert (out["n_windows"] >= 1).all()
    assert (out["n_windows"] >= 2).any(), "some drives must re-arm after a stop command"
    assert out["published"][:, 0].all()
    # the window the reference publishes: at most 149 samples, counts strictly inside (start, stop)
    assert out["n_samples"][:, 0].max() <= 149
    assert (np.diff(out["time_array"][0, 0, :out["n_samples"][0, 0]]) >= 1).all()


def test_stuck_samples_are_excluded():
    d = syn.drives(0, 64, T=300, stop_events=False)
    out = check(d)
    stuck_rows = (out["slip"] == 1.0).any(axis=1)
    assert stuck_rows.any()
    b = int(np.where(stuck_rows)[0][0])
    n = out["n_samples"][b, 0]
    assert n < 149 and not np.isin(1.0, out["slip_array"][b, 0, :n])


def test_never_driving_and_short_logs():
    d = syn.drives(0, 4, T=100)
    d["cmd"][:] = 0.0                                   # |cmd| <= 0.2: nothing is valid, nothing is armed
    out = check(d)
    assert (out["n_windows"] == 0).all() and (out["n_samples"] == 0).all()
    d = syn.drives(0, 4, T=60)                          # log ends before the window closes
    out = check(d)
    assert (out["n_windows"] == 0).all()


def test_window_below_fifteen_samples_is_not_published():
    d = syn.drives(3, 1, T=400, stop_events=False)
    lead = int(np.argmax(d["cmd"][0] > 0))
    d["cmd"][0, lead + 20:lead + 158] = 0.0             # driving pauses: only a few samples fall inside the window
    out = check(d)
    assert out["n_windows"][0] == 1 and 0 < out["n_samples"][0, 0] < 15 and out["published"][0, 0] == 0


def test_unexpected_stop_reinitialises():
    d = syn.drives(5, 1, T=700, stop_events=False)
    out = check(d)                                       # no stop command: after stop + 100 updates the recorder re-arms
    assert out["n_windows"][0] >= 2


def test_capacity_overflow_is_counted_not_stored():
    d = syn.drives(0, 2, T=400, stop_events=False)
    out = check(d, max_windows=1, cap=32)
    assert (out["n_samples"][:, 0] > 32).all()
