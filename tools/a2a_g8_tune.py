"""NVLink all-to-all at the full box: exchange schedules side by side (S = 256 MiB per pair)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe()
g = torch.cuda.device_count()
ords = list(range(g))
S = int(os.environ.get("A2A_S_MIB", "256")) << 20
out = []
def run(tag, variant=0, mode=0, cpp=None, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    try:
        r = p.nvlink_a2a(ords, bytes_per_pair=S, mode=mode, warmup=2, reps=8, variant=variant, ctas_per_peer=cpp)
        eg = [round(x) for x in r.egress_gbs[:g]]
        line = f"{tag} ms={r.ms_median:.4f} best={r.ms_best:.4f} per_dir_gbs={(g-1)*S/r.ms_median/1e6:.1f} best_gbs={(g-1)*S/r.ms_best/1e6:.1f} own={eg} verified={r.verified}"
    except Exception as e:
        line = f"{tag} ERR {str(e)[:160]}"
    for k in env:
        os.environ.pop(k, None)
    out.append(line); print(line, flush=True)
run("push_tma", 2)
run("push_stagger", 6)
run("push_stagger cta=296", 6, cpp=296)
run("push_stagger cta=74", 6, cpp=74)
run("push_tma cpp=42", 2, cpp=42)
run("push_tma cpp=10", 2, cpp=10)
run("pull_tma", 1)
run("mix90", 5, B200PROBE_A2A_MIX_PCT=90)
run("push_stagger SB=16384 NS=3", 6, B200PROBE_A2A_STAGE_BYTES=16384, B200PROBE_A2A_STAGES=3)
run("push_tma SB=16384 NS=3", 2, B200PROBE_A2A_STAGE_BYTES=16384, B200PROBE_A2A_STAGES=3)
run("push_direct", 3)
run("nccl", 0, mode=2)
run("push_tma again", 2)
run("push_stagger again", 6)
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/a2a_sched_g{g}.txt", "w").write("\n".join(out) + "\n")
