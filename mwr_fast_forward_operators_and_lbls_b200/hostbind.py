"""Host-side placement for the host entry point (`Plan.forward_host`): run the calling thread -- and
therefore first-touch its pinned buffers -- on the NUMA node the GPU's PCIe root hangs off.

With one process per GPU, eight ranks each move ~90 MB per 1.6 ms step between pinned host memory and
their GPU; if every rank's pages sit on one socket the step is bound by that socket's DRAM and by the
inter-socket link instead of by PCIe (end-to-end scaling 0.26 at 8 ranks).  No libnuma
dependency: Linux places freshly touched pages on the node of the touching CPU, so binding the CPU
affinity BEFORE the buffers are allocated is enough.
"""
import os


def gpu_numa_node(index):
    """NUMA node of CUDA device `index` from sysfs, or None when unknown (-1, no sysfs, no torch)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as fh:
            node = int(fh.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


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


def node_cpus(node):
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            return _parse_cpulist(fh.read())
    except Exception:
        return set()


def gpu_cpu_affinity(index):
    """CPUs NVML recommends for CUDA device `index` (its NUMA node's cores), or an empty set.  The
    fallback when sysfs does not expose numa_node (containers often show -1)."""
    try:
        import pynvml
        import torch
        pr = torch.cuda.get_device_properties(index)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id,
                                                                          pr.pci_device_id)).encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = set()
        for w, mask in enumerate(words):
            for b in range(64):
                if (int(mask) >> b) & 1:
                    cpus.add(64 * w + b)
        return cpus
    except Exception:
        return set()


def bind_to_gpu_node(index):
    """Restrict this process to the CPUs of the GPU's NUMA node (intersected with the CPUs it may use).
    Returns a dict describing what was done (for logs / the bench line); never raises."""
    info = {"gpu": int(index), "numa_node": None, "bound": False}
    try:
        node = gpu_numa_node(index)
        info["numa_node"] = node
        allowed = os.sched_getaffinity(0)
        if node is not None:
            want = node_cpus(node) & allowed
            info["source"] = "sysfs numa_node"
        else:
            want = gpu_cpu_affinity(index) & allowed
            info["source"] = "NVML cpu affinity"
            if not want:
                info["why"] = "GPU NUMA node unknown (sysfs and NVML)"
                return info
        info["cpus_allowed"] = len(allowed)
        if not want:
            info["why"] = "no allowed CPU near the GPU"
            return info
        if want == allowed:
            info["why"] = "already confined to the GPU's CPUs"
        os.sched_setaffinity(0, want)
        info["bound"] = True
        info["cpus_bound"] = len(want)
    except Exception as exc:  # noqa: BLE001
        info["why"] = repr(exc)
    return info
