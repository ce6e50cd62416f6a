"""Does pinned-buffer placement (NUMA node) change the PCIe copy bandwidth on this box?"""
import glob
import json
import os
import time

import torch


def cpus_of(node):
    txt = open('/sys/devices/system/node/node%d/cpulist' % node).read().strip()
    out = []
    for part in txt.split(','):
        if '-' in part:
            a, b = part.split('-')
            out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def main():
    res = {}
    nodes = sorted(int(p.rsplit('node', 1)[1]) for p in glob.glob('/sys/devices/system/node/node[0-9]*'))
    res['nodes'] = nodes
    res['affinity_at_start'] = len(os.sched_getaffinity(0))
    torch.cuda.init()
    prop = torch.cuda.get_device_properties(0)
    bus = '%04x:%02x:%02x.0' % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
    res['gpu_pci'] = bus
    try:
        res['gpu_numa_node'] = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read())
    except Exception as e:  # noqa: BLE001
        res['gpu_numa_node'] = repr(e)
    n = 112896000 // 4
    xd = torch.empty(n, device='cuda')
    all_cpus = sorted(os.sched_getaffinity(0))
    for node in nodes:
        cp = [c for c in cpus_of(node) if c in all_cpus]
        if not cp:
            continue
        os.sched_setaffinity(0, cp)
        time.sleep(0.05)
        xh = torch.empty(n, pin_memory=True)
        xh.fill_(1.0)
        for direction in ('h2d', 'd2h'):
            fn = (lambda: xd.copy_(xh, non_blocking=True)) if direction == 'h2d' else (lambda: xh.copy_(xd, non_blocking=True))
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            res['node%d_%s_GBs' % (node, direction)] = n * 4 * 10 / (time.perf_counter() - t0) / 1e9
        del xh
    os.sched_setaffinity(0, all_cpus)
    # several buffers allocated back to back without any binding
    bufs = [torch.empty(n, pin_memory=True) for _ in range(4)]
    for i, b in enumerate(bufs):
        b.fill_(1.0)
        xd.copy_(b, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            xd.copy_(b, non_blocking=True)
        torch.cuda.synchronize()
        res['unbound_buf%d_h2d_GBs' % i] = n * 4 * 10 / (time.perf_counter() - t0) / 1e9
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
