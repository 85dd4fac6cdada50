#!/usr/bin/env bash
# Box probe: facts about the gpurun B200 box that shape the design (SURVEY.md §7 "hard parts").
# Writes everything to gpurun_out/box_probe.txt. Never fails the call.
set +e
mkdir -p gpurun_out
out=gpurun_out/box_probe.txt
{
echo "== date"; date -u
echo "== nproc / cpu"; nproc; lscpu | grep -E 'Model name|Socket|Thread|Core|NUMA node\(s\)|^CPU\(s\)'
echo "== mem"; free -g | head -2
echo "== nvidia-smi -L"; nvidia-smi -L
echo "== nvidia-smi"; nvidia-smi
echo "== topo"; nvidia-smi topo -m
echo "== nvlink status"; nvidia-smi nvlink --status -i 0 | head -30
echo "== libs"; ldconfig -p | grep -E 'nvidia-ml|libcuda\.|libnccl'
echo "== toolchains"; which go ncu nsys nvcc compute-sanitizer
echo "== env"; env | grep -E 'CUDA|NVIDIA|NCCL'
echo "== python probe"
python - <<'EOF'
import time
t=time.time()
import torch
print("import torch s", round(time.time()-t,1))
print("cuda avail", torch.cuda.is_available(), "count", torch.cuda.device_count())
for i in range(torch.cuda.device_count()):
    p=torch.cuda.get_device_properties(i)
    print(i,p.name,p.total_memory,p.multi_processor_count,p.major,p.minor, getattr(p,'L2_cache_size',None))
n=torch.cuda.device_count()
for i in range(n):
    for j in range(n):
        if i!=j: print("p2p",i,j,torch.cuda.can_device_access_peer(i,j))
try:
    import pynvml
    t=time.time(); pynvml.nvmlInit(); print("nvmlInit s", round(time.time()-t,3))
    print("driver", pynvml.nvmlSystemGetDriverVersion(), "nvml", pynvml.nvmlSystemGetNVMLVersion())
    c=pynvml.nvmlDeviceGetCount(); print("nvml count", c)
    for i in range(c):
        h=pynvml.nvmlDeviceGetHandleByIndex(i)
        print(i, pynvml.nvmlDeviceGetUUID(h), pynvml.nvmlDeviceGetName(h), pynvml.nvmlDeviceGetMemoryInfo(h).total,
              pynvml.nvmlDeviceGetCudaComputeCapability(h))
        try: print(" mig", pynvml.nvmlDeviceGetMigMode(h))
        except Exception as e: print(" mig err", e)
        try: print(" numa", pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception as e: print(" numa err", e)
        try: print(" supported events", hex(pynvml.nvmlDeviceGetSupportedEventTypes(h)))
        except Exception as e: print(" events err", e)
        try:
            es=pynvml.nvmlEventSetCreate()
            pynvml.nvmlDeviceRegisterEvents(h, pynvml.nvmlEventTypeXidCriticalError|pynvml.nvmlEventTypeDoubleBitEccError|pynvml.nvmlEventTypeSingleBitEccError, es)
            t=time.time()
            try: d=pynvml.nvmlEventSetWait_v2(es, 10); print(" event", d.eventType, d.eventData)
            except pynvml.NVMLError as e: print(" wait ->", e, "in", round((time.time()-t)*1e3,2),"ms")
            pynvml.nvmlEventSetFree(es)
        except Exception as e: print(" eventset err", e)
        try:
            for l in range(18):
                print(" link",l,pynvml.nvmlDeviceGetNvLinkState(h,l), end=";")
            print()
        except Exception as e: print(" nvlink err", e)
except Exception as e:
    print("pynvml err", repr(e))
EOF
echo "== ncu permission check"
cat > /tmp/t.cu <<'EOF'
#include <cstdio>
__global__ void k(float* p){ p[threadIdx.x]=threadIdx.x; }
int main(){ float* p; cudaMalloc(&p,1024); k<<<1,32>>>(p); printf("rc %d\n",(int)cudaDeviceSynchronize());
 cudaDeviceProp pr; cudaGetDeviceProperties(&pr,0); printf("l2 %d smem optin %zu busw %d memclk %d sms %d\n", pr.l2CacheSize, pr.sharedMemPerBlockOptin, pr.memoryBusWidth, pr.memoryClockRate, pr.multiProcessorCount);
 return 0; }
EOF
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/t /tmp/t.cu && /tmp/t && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none /tmp/t 2>&1 | tail -15
echo "== done"
} > $out 2>&1
tail -5 $out
exit 0
