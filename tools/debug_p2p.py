import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe()
N = 1 << 20
a0 = torch.zeros(N, dtype=torch.uint8, device="cuda:0")
a1 = torch.zeros(N, dtype=torch.uint8, device="cuda:1")
print("can", torch.cuda.can_device_access_peer(0, 1), torch.cuda.can_device_access_peer(1, 0))
step = sys.argv[1] if len(sys.argv) > 1 else "lib"
if step == "torch":   # let torch enable peer access its own way first
    b = a0.to("cuda:1"); torch.cuda.synchronize(); print("torch p2p copy ok")
else:
    ords = (C.c_int * 2)(0, 1)
    print("enable rc", p.lib.b200probe_enable_peer_access(ords, 2))
torch.cuda.set_device(0)
st = torch.cuda.current_stream(0).cuda_stream
try:
    p.hbm_fill(0, a1.data_ptr(), N, 7, st, variant=1)      # dev0 kernel, direct STG to dev1 memory
    torch.cuda.synchronize(0); print("direct fill to peer ok", int(a1[:4].cpu().view(torch.int32)[0]))
    p.hbm_fill(0, a1.data_ptr(), N, 8, st)                 # TMA bulk store to peer
    torch.cuda.synchronize(0); print("tma fill to peer ok")
    p.hbm_copy(0, a1.data_ptr(), a0.data_ptr(), N, st)     # TMA bulk load from peer
    torch.cuda.synchronize(0); print("tma copy from peer ok", bool((a0.cpu() == a1.cpu()).all()))
    peers = (C.c_void_p * 2)(a0.data_ptr(), a1.data_ptr())
    rc = p.lib.b200probe_a2a_push(0, 0, 2, peers, 1 << 18, 0xB200, 3, st)
    torch.cuda.synchronize(0); print("a2a push rc", rc, "ok")
except Exception as e:
    print("FAIL", e)
