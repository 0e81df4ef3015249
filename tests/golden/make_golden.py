"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libvxref.so).

Run in the build container where /root/reference exists:
    make -C oracle ref && python tests/golden/make_golden.py
Each file holds the final voxel/link state of one case of tests/cases.py after its steps,
exactly as the reference produced it.  The oracle port must reproduce these bit for bit
(tests/test_oracle.py), which pins the oracle on machines without the reference sources.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
import parity  # noqa: E402
from voxelyze_b200 import capi  # noqa: E402


def main():
    ref = capi.load_reference()
    assert ref.backend == "reference"
    for c in cases.CASES:
        sc = c.make()
        sim, dt, div = parity.run(ref, sc, c.steps, program=c.program)
        snap = parity.snapshot(sim)
        # keep fixtures small: float32 fields and flags in full, double fields in full for small cases
        out = {k: v for k, v in snap.items()}
        out["dt"] = np.float32(dt)
        out["diverged"] = np.int32(-1 if div is None else div)
        if sc.collisions:
            out["pairs"] = sim.collision_pairs()
        np.savez_compressed(os.path.join(HERE, c.name + ".npz"), **out)
        print(f"{c.name:28s} nvox={sim.n_voxels:5d} nlink={sim.n_links:5d} dt={dt:.9g} div={div}")
        sim.close()


if __name__ == "__main__":
    main()
