"""NUMA placement of a rank's host side: bind the process to the CPU cores of the NUMA node its GPU hangs off, so that
the pinned staging buffers (allocated afterwards; first touch / cudaHostAlloc is node-local under the default policy)
and the threads that fill them sit next to the GPU's PCIe root port.  With one process per GPU (torchrun) all ranks
otherwise start on the same node and every device-to-host byte of the far GPUs crosses the socket interconnect.
Pure /sys + sched_setaffinity (no numactl / libnuma in this image); a no-op where the topology is not exposed."""
import glob
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index):
    """NUMA node of CUDA device `device_index` (PCI bus id from torch / NVML -> /sys/bus/pci/devices/*/numa_node); -1 if unknown."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        try:
            import pynvml
            pynvml.nvmlInit()
            bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
            bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()[-12:]
        except Exception:
            return -1
    path = f"/sys/bus/pci/devices/{bdf}/numa_node"
    try:
        with open(path) as f:
            return int(f.read().strip())
    except Exception:
        return -1


def bind_to_gpu_node(device_index):
    """Restrict this process to the cores of its GPU's NUMA node.  Returns a dict describing what was done."""
    nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
    node = gpu_numa_node(device_index)
    info = {"gpu": device_index, "node": node, "nodes": len(nodes), "bound": False}
    if node < 0 or len(nodes) < 2:
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as e:
        info["error"] = str(e)
    return info
