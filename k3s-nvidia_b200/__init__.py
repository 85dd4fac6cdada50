"""b200-gpu-health-probe: a B200-native active GPU health-probe path behind the NVIDIA
k8s-device-plugin surface that UntouchedWagons/K3S-NVidia installs (/root/reference/README.md:116,
configured by /root/reference/values.yaml:1-18).

Layout (DESIGN.md has the full map):
  csrc/            hand-written sm_100a CUDA kernels + the C ABI  -> libb200probe.so
  _lib.py          ctypes binding of include/b200probe.h (fails loudly when the .so is missing)
  probe.py         Python face of the C ABI (enumerate / passive health / HBM / NVLink / GEMM)
  config.py        the plugin config document of values.yaml:8-18 and time-slicing replicas
  api.py           kubelet device-plugin API v1beta1 messages + gRPC stubs (runtime descriptors)
  plugin.py        Register / ListAndWatch / Allocate / GetPreferredAllocation server
  labels.py        NFD features.d label writer (probe results -> node labels)
"""
__version__ = "0.1.0"
