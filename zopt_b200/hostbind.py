"""
Host-side placement for the end-to-end path (one process per GPU, only a final gather).

The solvers themselves never touch host memory; what a caller moves per step is its x0 batch in and the plans out, over
PCIe.  With eight ranks on a two-socket host those copies run at the link rate only if every rank's pinned buffers live
on the NUMA node its GPU hangs off, which is decided by where the allocating thread runs.  `bind_to_gpu(device)` pins
the calling process to the CPUs NVML reports as local to the device (before any pinned allocation); everything is
best-effort and reports what it did.  No torch/CUDA arithmetic here, and nothing imports the CUDA library.
"""
import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(s):
    cpus = set()
    for part in (s or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def pci_bus_id(device_index):
    """'dddddddd:bb:dd.0' of a CUDA device, from torch's device properties (CUDA_VISIBLE_DEVICES-safe)."""
    import torch
    p = torch.cuda.get_device_properties(device_index)
    return f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"


def gpu_numa_node(device_index):
    """NUMA node of the GPU's PCIe slot (sysfs), or None"""
    bid = pci_bus_id(device_index)
    for cand in (bid, bid[4:]):  # sysfs uses a 4-digit domain
        v = _read(f"/sys/bus/pci/devices/{cand.lower()}/numa_node")
        if v is not None:
            try:
                return int(v)
            except ValueError:
                return None
    return None


def bind_to_gpu(device_index, enable=True):
    """Restrict this process to the CPUs local to the GPU.  Returns a dict describing the outcome
    ({"bound": bool, "numa_node": int|None, "cpus": int, "how": str})."""
    info = {"bound": False, "numa_node": None, "cpus": len(os.sched_getaffinity(0)), "how": "unbound"}
    if not enable:
        return info
    try:
        info["numa_node"] = gpu_numa_node(device_index)
    except Exception:
        pass
    allowed = os.sched_getaffinity(0)
    # 1. sysfs: CPUs of the GPU's NUMA node
    node = info["numa_node"]
    if node is not None and node >= 0:
        cpus = _parse_cpulist(_read(f"/sys/devices/system/node/node{node}/cpulist")) & allowed
        if cpus and cpus != allowed:
            try:
                os.sched_setaffinity(0, cpus)
                info.update(bound=True, cpus=len(cpus), how=f"sysfs numa node {node}")
                return info
            except OSError:
                pass
        elif cpus:
            info["how"] = "single NUMA node visible (nothing to bind)"
            return info
    # 2. NVML's ideal CPU set for the device
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id(device_index).encode())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        now = os.sched_getaffinity(0)
        if now != allowed:
            info.update(bound=True, cpus=len(now), how="nvmlDeviceSetCpuAffinity")
    except Exception as e:  # containers without NVML access, cpuset restrictions, ...
        info["how"] = f"unbound ({type(e).__name__})"
    return info


def host_topology():
    """Small description of the host for the bench record: NUMA nodes, CPUs per node, CPUs this process may use."""
    nodes = {}
    try:
        for d in sorted(os.listdir("/sys/devices/system/node")):
            if d.startswith("node") and d[4:].isdigit():
                nodes[int(d[4:])] = len(_parse_cpulist(_read(f"/sys/devices/system/node/{d}/cpulist")))
    except OSError:
        pass
    return {"numa_nodes": len(nodes), "cpus_per_node": nodes, "cpus_allowed": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count()}
