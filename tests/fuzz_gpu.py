"""GPU counterpart of tests/fuzz_loader.py: damaged fixture images through the C ABI of the product on cuda:0. Every
image must be rejected or load into an index on which every kernel family runs to completion without a CUDA error.
Usage (GPU box, under `timeout`): python tests/fuzz_gpu.py <seed> <cases>."""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import gbwt_rs_b200 as gb  # noqa: E402

FIXTURES = ["example.gbwt", "with-empty.gbwt", "translation.gbz", "example-v1.gbz", "example.gbz", "translation-v1.gbz"]


def main(seed: int, cases: int) -> None:
    rng = random.Random(seed)
    loaded = rejected = 0
    for case in range(cases):
        raw = bytearray(open(os.path.join(HERE, "golden", rng.choice(FIXTURES)), "rb").read())
        for _ in range(rng.choice([1, 1, 2, 4, 16])):
            i = rng.randrange(len(raw))
            raw[i] = rng.randrange(256) if rng.random() < 0.5 else raw[i] ^ (1 << rng.randrange(8))
        if rng.random() < 0.2:
            raw = raw[:rng.randrange(len(raw))]
        try:
            e = gb.GBWT.from_bytes(bytes(raw), layout=rng.choice(["auto", "runs"]))
        except (IOError, OSError, gb.GBWTError):
            rejected += 1
            continue
        loaded += 1
        n = e.alphabet_size() + 8
        nodes = np.array([rng.randrange(0, n) for _ in range(64)], dtype=np.uint64)
        st = e.find(nodes)
        e.extend(st, nodes[::-1].copy())
        e.find_extend(np.array([[rng.randrange(0, n) for _ in range(5)] for _ in range(64)], dtype=np.uint64))
        if e.is_bidirectional():
            bd = e.bd_find(nodes)
            e.extend_forward(bd, nodes[::-1].copy())
            e.extend_backward(bd, nodes[::-1].copy())
            e.follow(bd[:16])
            e.follow(bd[:16], backward=True)
        pos = np.zeros(64, dtype=[("node", "<u8"), ("offset", "<u8")])
        pos["node"] = nodes
        pos["offset"] = [rng.randrange(0, 12) for _ in range(64)]
        e.forward(pos)
        if e.is_bidirectional():
            e.backward(pos)
        ids = np.arange(0, 8, dtype=np.uint64)
        e.extract(ids)
        e.extract(ids)          # lengths known: two-ended where long enough
        if e.has_graph():
            e.extract_dna(ids)
            e.extract_dna(ids)
            e.node_sequences(np.arange(0, 30, dtype=np.uint64))
        e.serialize()
        e.close()
    print(f"loaded {loaded} rejected {rejected}")


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]))
