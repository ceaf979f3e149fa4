"""Prints CU_DEVICE_ATTRIBUTE_TENSOR_MAP_ACCESS_SUPPORTED (127) and a few neighbours for device 0 (TEST TOOL)."""
import ctypes as C

cu = C.CDLL("libcuda.so.1")
assert cu.cuInit(0) == 0
dev = C.c_int()
assert cu.cuDeviceGet(C.byref(dev), 0) == 0
for name, attr in (("TENSOR_MAP_ACCESS_SUPPORTED", 127), ("CLUSTER_LAUNCH", 120), ("COMPUTE_CAPABILITY_MAJOR", 75), ("COMPUTE_CAPABILITY_MINOR", 76),
                   ("UNIFIED_FUNCTION_POINTERS", 129), ("MPS_ENABLED", 133), ("HOST_NUMA_ID", 134), ("PAGEABLE_MEMORY_ACCESS", 88)):
    v = C.c_int(-1)
    rc = cu.cuDeviceGetAttribute(C.byref(v), attr, dev)
    print(f"{name} ({attr}): rc {rc} value {v.value}")
