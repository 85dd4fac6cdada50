"""One active-probe round of the Python runner, standalone (what `python -m k3s_nvidia_b200.plugin` runs every interval), twice:
labels that are not "true", per-GPU egress, wall time."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from k3s_nvidia_b200.labels import ActiveProbeRunner, PREFIX
from k3s_nvidia_b200.probe import Probe
os.environ["B200PROBE_IGNORE_TENANTS"] = "1"
p = Probe()
with tempfile.TemporaryDirectory() as d:
    r = ActiveProbeRunner(p, features_dir=d, keep_arenas=("--keep" in sys.argv))
    for i in range(2):
        t0 = time.perf_counter()
        lab = r.run_once()
        print(f"round {i}: {time.perf_counter() - t0:.3f} s gate={lab.get(PREFIX + 'healthy')} egress={[v for k, v in sorted(lab.items()) if k.endswith('nvlink-egress-gbs')]} "
              f"not_true={[k[len(PREFIX):] for k, v in lab.items() if v == 'false']}", flush=True)
