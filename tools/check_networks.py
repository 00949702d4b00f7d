#!/usr/bin/env python
"""0-1 principle check of the compare-exchange networks in lidardetection_b200/csrc/lg_geom.cuh (sort8, sort16):
a network sorts every input iff it sorts every 0/1 input.  Also run by tests/test_host_api.py."""
import itertools
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def networks():
    src = open(os.path.join(ROOT, "lidardetection_b200", "csrc", "lg_geom.cuh")).read()
    out = {}
    for m in re.finditer(r"void (sort\d+)\(uint32_t \(&k\)\[(\d+)\]\) \{(.*?)\n\}", src, re.S):
        out[m.group(1)] = (int(m.group(2)), [(int(a), int(b)) for a, b in re.findall(r"cex\(k\[(\d+)\], k\[(\d+)\]\)", m.group(3))])
    return out


def sorts(n, net):
    for bits in itertools.product((0, 1), repeat=n):
        a = list(bits)
        for i, j in net:
            if a[i] > a[j]:
                a[i], a[j] = a[j], a[i]
        if any(a[i] > a[i + 1] for i in range(n - 1)):
            return False
    return True


if __name__ == "__main__":
    for name, (n, net) in sorted(networks().items()):
        print(name, n, "inputs,", len(net), "compare-exchanges:", "OK" if sorts(n, net) else "BROKEN")
