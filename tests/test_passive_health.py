"""Passive health: libb200probe.so vs the oracle twin on the same scripted NVML event streams.
Verdict parity is bit-exact: same device order, same UUIDs, same Healthy/Unhealthy strings."""
import ctypes as C
import os

import pytest

import _oracle

XID, DBE, SBE, WAIT_ERR, XID_NOUUID, XID_UNKNOWN = 0, 1, 2, 3, 4, 5

SCENARIOS = {
    "no-events": (4, None, {}, []),
    "critical-xid-one-gpu": (4, None, {}, [(XID, 2, 79)]),
    "application-xids-skipped": (4, None, {}, [(XID, 0, 13), (XID, 1, 31), (XID, 2, 43), (XID, 3, 45), (XID, 0, 68), (XID, 1, 109)]),
    "ecc-events-are-not-xid": (2, None, {}, [(DBE, 0, 0), (SBE, 1, 0)]),
    "extra-skip-list": (4, "48, 79 ,bogus,,", {}, [(XID, 1, 48), (XID, 2, 79), (XID, 3, 62)]),
    "disabled-all": (4, "all", {}, [(XID, 0, 79)]),
    "disabled-xids": (4, "XIDs", {}, [(XID, 0, 79)]),
    "wait-error-marks-all": (3, None, {}, [(WAIT_ERR, 0, 15)]),
    "uuid-unreadable-marks-all": (3, None, {}, [(XID_NOUUID, 0, 79)]),
    "unknown-device-ignored": (3, None, {}, [(XID_UNKNOWN, 0, 79)]),
    "register-not-supported": (4, None, {"MOCK_NVML_NO_EVENTS": "1,3"}, []),
    "supported-query-fails": (2, None, {"MOCK_NVML_EVENTS_QUERY_FAIL": "0"}, []),
    "sticky-unhealthy": (2, None, {}, [(XID, 0, 79), (XID, 0, 13), (SBE, 0, 0)]),
    "eight-gpus-two-fail": (8, "", {}, [(XID, 7, 119), (XID, 0, 94), (XID, 3, 45)]),
}


def _mock():
    m = C.CDLL(_oracle.MOCK_NVML)
    m.mock_nvml_push.argtypes = [C.c_int, C.c_int, C.c_ulonglong]
    m.mock_nvml_registered.restype = C.c_ulonglong
    return m


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_product_matches_oracle(name, monkeypatch):
    from k3s_nvidia_b200.probe import Probe

    ndev, disable, env, events = SCENARIOS[name]
    monkeypatch.setenv("MOCK_NVML_DEVICES", str(ndev))
    for k in ("MOCK_NVML_NO_EVENTS", "MOCK_NVML_EVENTS_QUERY_FAIL", "MOCK_NVML_UUID_FAIL"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    mock = _mock()

    # --- product
    p = Probe(_oracle.MOCK_NVML)
    try:
        infos = [p.device_info(i) for i in range(p.device_count())]
        mask = p.health_open(disable if disable is not None else "")
        for e in events:
            mock.mock_nvml_push(*e)
        for _ in range(len(events) + 2):
            p.health_wait(1)
        mask = p.health_mask()
        product = [(d.index, d.uuid, d.name, d.mem_total, d.cc, "Unhealthy" if (mask >> d.index) & 1 else "Healthy") for d in infos]
    finally:
        p.health_close()
        p.close()

    # --- oracle twin (fresh NVML session, same script)
    o = _oracle.load()
    assert o.oracle_ph_open(_oracle.MOCK_NVML.encode(), disable.encode() if disable is not None else None) == 0
    try:
        for e in events:
            mock.mock_nvml_push(*e)
        for _ in range(len(events) + 2):
            o.oracle_ph_poll(1)
        oracle = _oracle.verdicts(o)
    finally:
        o.oracle_ph_close()

    assert product == oracle

    # scenario-specific expectations (so both being wrong the same way is caught too)
    unhealthy = {i for (i, *_r, h) in product if h == "Unhealthy"}
    expect = {
        "no-events": set(), "critical-xid-one-gpu": {2}, "application-xids-skipped": set(), "ecc-events-are-not-xid": set(),
        "extra-skip-list": {3}, "disabled-all": set(), "disabled-xids": set(), "wait-error-marks-all": {0, 1, 2},
        "uuid-unreadable-marks-all": {0, 1, 2}, "unknown-device-ignored": set(), "register-not-supported": {1, 3},
        "supported-query-fails": {0}, "sticky-unhealthy": {0}, "eight-gpus-two-fail": {0, 7},
    }[name]
    assert unhealthy == expect


def test_registers_xid_and_ecc_events_masked_by_supported(monkeypatch):
    from k3s_nvidia_b200.probe import Probe

    monkeypatch.setenv("MOCK_NVML_DEVICES", "2")
    mock = _mock()
    p = Probe(_oracle.MOCK_NVML)
    try:
        assert p.health_open("") == 0
        assert mock.mock_nvml_registered(0) == 0x8 | 0x2 | 0x1
    finally:
        p.health_close()
        p.close()


def test_event_details_and_timeout(monkeypatch):
    from k3s_nvidia_b200.probe import Probe

    monkeypatch.setenv("MOCK_NVML_DEVICES", "2")
    mock = _mock()
    p = Probe(_oracle.MOCK_NVML)
    try:
        p.health_open("")
        ev = p.health_wait(1)
        assert ev.timed_out and ev.newly_unhealthy == 0
        mock.mock_nvml_push(XID, 1, 79)
        ev = p.health_wait(1)
        assert (ev.event_type, ev.event_data, ev.device_index, ev.skipped, ev.newly_unhealthy) == (8, 79, 1, False, 2)
        mock.mock_nvml_push(XID, 1, 79)
        ev = p.health_wait(1)
        assert ev.newly_unhealthy == 0            # already unhealthy: no second transition
        assert p.passive_health(1) == 2           # one-shot form returns the sticky mask
    finally:
        p.health_close()
        p.close()


def test_wait_before_open_is_a_state_error(monkeypatch):
    from k3s_nvidia_b200.probe import Probe, ProbeError

    monkeypatch.setenv("MOCK_NVML_DEVICES", "1")
    p = Probe(_oracle.MOCK_NVML)
    try:
        with pytest.raises(ProbeError) as e:
            p.health_wait(1)
        assert e.value.rc == -11
    finally:
        p.close()


def test_close_waits_for_a_thread_blocked_in_the_event_wait(monkeypatch):
    """b200probe_health_wait drops the library lock around nvmlEventSetWait_v2; a concurrent health_close (or
    shutdown) used to free the event set under the blocked waiter.  Close now waits until the waiter has left, and the waiter
    that wakes up on a closed set reports a timeout instead of touching it."""
    import threading
    import time

    from k3s_nvidia_b200.probe import Probe

    monkeypatch.setenv("MOCK_NVML_DEVICES", "2")
    monkeypatch.setenv("MOCK_NVML_WAIT_FULL", "1")
    p = Probe(_oracle.MOCK_NVML)
    try:
        p.health_open("")
        out = {}

        def waiter():
            t0 = time.perf_counter()
            out["ev"] = p.health_wait(400)
            out["dt"] = time.perf_counter() - t0

        import ctypes as C

        mock = C.CDLL(_oracle.MOCK_NVML)               # same handle the library dlopen()ed: its wait counter tells when the waiter is inside
        before = mock.mock_nvml_wait_calls()
        th = threading.Thread(target=waiter)
        th.start()
        t_end = time.time() + 5
        while mock.mock_nvml_wait_calls() == before and time.time() < t_end:
            time.sleep(0.005)
        assert mock.mock_nvml_wait_calls() > before, "the waiter never reached nvmlEventSetWait_v2"
        time.sleep(0.02)
        t0 = time.perf_counter()
        p.health_close()                                    # must not return while the waiter is inside NVML
        closed_after = time.perf_counter() - t0
        th.join(5)
        assert not th.is_alive()
        assert closed_after > 0.25, f"close returned after {closed_after:.3f} s with a waiter still blocked"
        assert out["dt"] > 0.35 and out["ev"].timed_out and out["ev"].newly_unhealthy == 0
        assert p.health_open("") == 0                       # the library is in a clean state: open / wait work again
        assert p.health_wait(1).timed_out
    finally:
        p.close()
