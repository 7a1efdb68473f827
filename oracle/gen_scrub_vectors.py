#!/usr/bin/env python3
"""Writes tests/golden/scrub_vectors.json: known-answer vectors for the scrub contract.

ORACLE / TEST INFRASTRUCTURE.  The reference has no scrub, so these
vectors are constructed from the contract: k in
{0, 1, 7, 4096} non-zero bytes at numpy.random.default_rng(1234) offsets, always
including offset 0, the last byte and one non-16-aligned offset; a 0xA5 poison; and
one pinned instance of the seeded sparse pattern.  Expected counts come from numpy
(np.count_nonzero), i.e. independently of both the C oracle and the CUDA kernels.
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import scrub_oracle as SO  # noqa: E402


def injection(nbytes, k, rng):
    offs = set()
    if k >= 1:
        offs.add(0)
    if k >= 2:
        offs.add(nbytes - 1)
    if k >= 3:
        offs.add(min(nbytes - 2, 16 * (nbytes // 32) + 3))  # not 16-aligned
    while len(offs) < k:
        offs.add(int(rng.integers(0, nbytes)))
    return [[o, int(rng.integers(1, 256))] for o in sorted(offs)]


def main():
    rng = np.random.default_rng(1234)
    vectors = []
    for nbytes in (1, 15, 16, 17, 4099, (1 << 20) + 5):
        for k in (0, 1, 7, 4096):
            if k > nbytes:
                continue
            poke = injection(nbytes, k, rng)
            buf = np.zeros(nbytes, dtype=np.uint8)
            for o, v in poke:
                buf[o] = v
            vectors.append(dict(name=f"zeros_{nbytes}_k{k}", nbytes=nbytes, fill=None, poke=poke,
                                nonzero=int(np.count_nonzero(buf))))
    for nbytes in (1, 4099):
        vectors.append(dict(name=f"poison_a5_{nbytes}", nbytes=nbytes, fill=0xA5, poke=[], nonzero=nbytes))
    vectors.append(dict(name="poison_with_zero_holes", nbytes=1000, fill=0xA5, poke=[[0, 0], [499, 0], [999, 0]],
                        nonzero=997))
    nbytes, seed, word0 = 65536 + 4, 1234, 7
    pat = SO.pattern_np(nbytes, seed, word0)
    doc = dict(generated_by="oracle/gen_scrub_vectors.py", vectors=vectors,
               pattern=dict(nbytes=nbytes, seed=seed, word_index0=word0, nonzero=int(np.count_nonzero(pat)),
                            first_bytes=pat[:64].tolist(), sha256=hashlib.sha256(pat.tobytes()).hexdigest()))
    path = ROOT / "tests" / "golden" / "scrub_vectors.json"
    path.write_text(json.dumps(doc, separators=(",", ":")) + "\n")
    print("wrote", path, len(vectors), "vectors")


if __name__ == "__main__":
    main()
