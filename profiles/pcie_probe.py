#!/usr/bin/env python
"""PCIe probe for the e2e path: pinned H2D / D2H bandwidth alone and concurrently, per NUMA node of the host buffers.

Measurement tooling only (uses torch for brevity; the product does not).  Run on the GPU box:
    python profiles/pcie_probe.py
"""
import glob
import os
import time

import torch


def cpus_of(node):
    txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = []
    for part in txt.split(","):
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def main():
    dev = torch.device("cuda:0")
    prop = torch.cuda.get_device_properties(0)
    bdf = None
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        if isinstance(bdf, bytes):
            bdf = bdf.decode()
        bdf = bdf.lower()[-12:]
        print("gpu", prop.name, "bdf", bdf, "numa_node", open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
    except Exception as e:  # noqa: BLE001
        print("nvml/sysfs:", e)
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    print("numa nodes", nodes, "cpus", os.cpu_count())
    n = 288_000_000
    m = 216_836_800
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(m, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for node in nodes + [None]:
        if node is not None:
            os.sched_setaffinity(0, cpus_of(node))
        else:
            os.sched_setaffinity(0, range(os.cpu_count()))
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(m, dtype=torch.uint8).pin_memory()
        h_in.fill_(1)
        h_out.fill_(1)
        res = {}
        for name in ("h2d", "d2h", "both"):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                if name in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if name in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 10
            res[name] = dt
        print(f"node {node}: h2d {n / res['h2d'] / 1e9:.1f} GB/s, d2h {m / res['d2h'] / 1e9:.1f} GB/s, "
              f"both {1e3 * res['both']:.2f} ms/step (h2d {n / res['both'] / 1e9:.1f} + d2h {m / res['both'] / 1e9:.1f} GB/s)")
        del h_in, h_out


if __name__ == "__main__":
    main()
