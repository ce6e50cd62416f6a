"""Host-memory placement for the end-to-end (host buffer in, host buffer out) path.

On a two-socket host the PCIe root of a GPU hangs off ONE socket; page-locked staging buffers that
live on the other socket cross the inter-socket link on every copy, and with one process per GPU all
eight processes then compete for it (round 1: end-to-end throughput at 8 GPUs was 0.67 of 8 x the
1-GPU figure while the device-resident metric scaled 0.99).  ``bind_to_gpu_node`` pins the calling
process to the CPUs of the GPU's NUMA node BEFORE the pinned pools are created, so first-touch places
them on the local node.  Pure host logic (sysfs + sched_setaffinity): no effect on results.
"""
from __future__ import annotations

import glob
import os


def _parse_cpulist(txt: str):
    out = []
    for part in txt.strip().split(','):
        if '-' in part:
            a, b = part.split('-')
            out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def numa_nodes():
    return sorted(int(p.rsplit('node', 1)[1]) for p in glob.glob('/sys/devices/system/node/node[0-9]*'))


def cpus_of_node(node: int):
    with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
        return _parse_cpulist(f.read())


def gpu_numa_node(device_index: int):
    """NUMA node of a CUDA device from its PCI address (None when sysfs has no answer)."""
    import torch
    prop = torch.cuda.get_device_properties(device_index)
    bus = '%04x:%02x:%02x.0' % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
    try:
        with open('/sys/bus/pci/devices/%s/numa_node' % bus) as f:
            node = int(f.read())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def bind_to_gpu_node(device_index: int):
    """Restrict the process to the CPUs of the device's NUMA node.  Returns a dict describing what was
    done ({'node': n, 'cpus': k, 'previous': [...]}); 'node' is None when nothing was changed."""
    prev = sorted(os.sched_getaffinity(0))
    info = {'node': None, 'cpus': len(prev), 'previous': prev}
    try:
        node = gpu_numa_node(device_index)
        if node is None or len(numa_nodes()) < 2:
            return info
        cpus = [c for c in cpus_of_node(node) if c in prev]
        if not cpus:
            return info
        os.sched_setaffinity(0, cpus)
        info.update(node=node, cpus=len(cpus))
    except OSError:
        pass
    return info


def restore_affinity(info):
    """Undo ``bind_to_gpu_node`` (e.g. before a CPU baseline that should see every core)."""
    try:
        if info and info.get('previous'):
            os.sched_setaffinity(0, info['previous'])
    except OSError:
        pass
