"""Allocate / GetPreferredAllocation round-trip latency of both hosts over the unix socket (mock NVML, grpcio client):
SURVEY.md §8a rows a4/a8 say "one unix-socket gRPC RTT (sub-ms)".  CPU only.  Usage: python tools/rpc_latency.py [calls]"""
import os, statistics, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["PYTHONPATH"] = ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")
import test_native_plugin as T
from k3s_nvidia_b200 import api

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
for name, cls in (("native (host/cpp)", T.Daemon), ("python (grpcio server)", T.PyDaemon)):
    d = tempfile.mkdtemp()
    k = T.FakeKubelet(d); k.start()
    dm = cls(d, T.VALUES.raw_configs["default"])
    assert k.event.wait(30) and dm.wait_serving(30), dm.logtext()
    ids = [f"{T.U0}::{r}" for r in range(4)] + [f"{T.U1}::{r}" for r in range(4)]
    with k.plugin_channel() as ch:
        stub = api.DevicePluginStub(ch)
        req_a = api.AllocateRequest(container_requests=[api.ContainerAllocateRequest(devices_ids=[ids[2]])])
        req_p = api.PreferredAllocationRequest(container_requests=[api.ContainerPreferredAllocationRequest(available_deviceIDs=ids, allocation_size=2)])
        for label, fn, req in (("Allocate", stub.Allocate, req_a), ("GetPreferredAllocation", stub.GetPreferredAllocation, req_p)):
            for _ in range(200):
                fn(req)
            ts = []
            for _ in range(N):
                t0 = time.perf_counter(); fn(req); ts.append((time.perf_counter() - t0) * 1e6)
            ts.sort()
            print(f"{name:24s} {label:24s} p50 {statistics.median(ts):7.1f} us   p99 {ts[int(0.99 * N)]:7.1f} us   ({N} calls, client = grpcio)")
    dm.stop(); k.stop()
