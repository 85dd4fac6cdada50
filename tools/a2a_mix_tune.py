"""NVLink a2a: MIX_TMA pushed-share sweep + ring geometry, single-process probe over all visible GPUs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe()
g = torch.cuda.device_count()
ords = list(range(g))
S = int(os.environ.get("A2A_S_MIB", "256")) << 20
out = []
def run(tag, variant, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    try:
        r = p.nvlink_a2a(ords, bytes_per_pair=S, mode=0, warmup=2, reps=8, variant=variant, ctas_per_peer=env.get("_CPP") or None)
        line = f"{tag} ms={r.ms_median:.4f} best={r.ms_best:.4f} per_dir_gbs={(g-1)*S/r.ms_median/1e6:.1f} egress_min={min(r.egress_gbs[:g]):.1f} egress_max={max(r.egress_gbs[:g]):.1f} verified={r.verified}"
    except Exception as e:
        line = f"{tag} ERR {str(e)[:160]}"
    for k in env:
        os.environ.pop(k, None)
    out.append(line); print(line, flush=True)
run("push_tma", 2)
run("pull_tma", 1)
for pct in (10, 25, 40, 50, 60, 75, 90):
    run(f"mix pct={pct}", 5, B200PROBE_A2A_MIX_PCT=pct)
for sb, ns in ((4096, 8), (16384, 2), (16384, 3), (32768, 2), (8192, 6), (2048, 8)):
    run(f"push_tma SB={sb} NS={ns}", 2, B200PROBE_A2A_STAGE_BYTES=sb, B200PROBE_A2A_STAGES=ns)
    run(f"mix50 SB={sb} NS={ns}", 5, B200PROBE_A2A_STAGE_BYTES=sb, B200PROBE_A2A_STAGES=ns)
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/a2a_mix_g{g}.txt", "w").write("\n".join(out) + "\n")
