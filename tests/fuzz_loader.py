"""Mutation fuzzer for the product's host side (loader, layout builder, writer) and its one-lane query code, run on
the CPU through tests/hostsim: damaged images must be rejected or load into something every query can run on without
crashing or looping for ever. Usage: python tests/fuzz_loader.py <seed> <cases>. Prints 'loaded N rejected M'.
(Found so far: absurd declared sizes reaching an allocation, a walk that never ends on an index with a cycle, and a
dense body of zero blocks for a record with edges but no runs.)"""
import faulthandler
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from hostsim_build import HostSim  # noqa: E402

FIXTURES = ["example.gbwt", "with-empty.gbwt", "translation.gbz", "example-v1.gbz", "example.gbz", "translation-v1.gbz"]


def main(seed: int, cases: int) -> None:
    faulthandler.enable()
    rng = random.Random(seed)
    loaded = rejected = 0
    for _ in range(cases):
        faulthandler.cancel_dump_traceback_later()
        faulthandler.dump_traceback_later(30, exit=True)  # a hang is a failure too
        raw = bytearray(open(os.path.join(HERE, "golden", rng.choice(FIXTURES)), "rb").read())
        for _ in range(rng.choice([1, 1, 2, 4, 16])):
            i = rng.randrange(len(raw))
            raw[i] = rng.randrange(256) if rng.random() < 0.5 else raw[i] ^ (1 << rng.randrange(8))
        if rng.random() < 0.2:
            raw = raw[:rng.randrange(len(raw))]
        try:
            h = HostSim(bytes(raw), rng.choice([0, 1]))
        except (IOError, OSError):
            rejected += 1
            continue
        loaded += 1
        n = h.records()
        nodes = np.array([rng.randrange(0, n + 40) for _ in range(64)], dtype=np.uint64)
        st = h.find(nodes)
        h.extend(st, nodes[::-1].copy())
        h.find_extend(np.array([[rng.randrange(0, n + 40) for _ in range(5)] for _ in range(64)], dtype=np.uint64))
        bd = h.bd_find(nodes)
        h.extend_forward(bd, nodes[::-1].copy())
        h.extend_backward(bd, nodes[::-1].copy())
        h.follow(bd[:16])
        h.follow(bd[:16], backward=True)
        pos = np.zeros(64, dtype=[("node", "<u8"), ("offset", "<u8")])
        pos["node"] = nodes
        pos["offset"] = [rng.randrange(0, 12) for _ in range(64)]
        h.forward(pos)
        h.backward(pos)
        h.sequence_lengths(np.arange(0, 8, dtype=np.uint64))
        h.extract(np.arange(0, 6, dtype=np.uint64))
        h.labels()
        h.serialize()
    faulthandler.cancel_dump_traceback_later()
    print(f"loaded {loaded} rejected {rejected}")


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]))
