"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref, -msse2 build).

Run in the dev container (needs /root/reference to have been compiled by oracle/Makefile):
    python tests/golden/make_golden.py
The outputs are what QCSim's own headers produce for seeded inputs; the inputs are rebuilt
from the JSON spec by tests/golden_util.py, so fixtures stay small.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import golden_util  # noqa: E402
import oracle  # noqa: E402

SPECS = {
    "all_gates_n6": {"n": 6, "circuit": "all_gates", "start": "random", "state_seed": 5,
                     "measure": {"count": 16, "seed": 3, "ranges": [[0, 0], [5, 5], [1, 3]], "collapse_range": [2, 4]}},
    "all_gates_n3": {"n": 3, "circuit": "all_gates", "start": "random", "state_seed": 6},
    "qft_n10": {"n": 10, "circuit": "qft", "start": "random", "state_seed": 7},
    "iqft_n10_noswap": {"n": 10, "circuit": "qft", "inverse": True, "swap": False, "start": "random", "state_seed": 8},
    "qft_sub_n9": {"n": 9, "circuit": "qft", "sq": 2, "eq": 6, "start": "random", "state_seed": 9},
    "random_n12_l5": {"n": 12, "circuit": "random", "layers": 5, "start": "zero",
                      "measure": {"count": 32, "seed": 42, "ranges": [[0, 0], [11, 11], [3, 8]], "collapse_range": [4, 9]}},
    "random_n16_l8": {"n": 16, "circuit": "random", "layers": 8, "start": "zero", "stride": 61,
                      "measure": {"count": 8, "seed": 1, "ranges": [[15, 15]], "collapse_range": [0, 0]}},
    "grover_n5": {"n": 9, "circuit": "grover", "n_search": 5, "marked": 22, "start": "zero",
                  "measure": {"count": 8, "seed": 2, "ranges": [[0, 4]]}},
    # BASELINE config 1: 20-qubit QFT + IQFT + MeasureAll
    "qft20_basis": {"n": 20, "circuit": "qft", "start": "basis", "basis": 0x5A5A5, "stride": 4099,
                    "measure": {"count": 8, "seed": 42, "ranges": [[0, 9]]}},
    "qft_iqft20_random": {"n": 20, "circuit": "qft_iqft", "start": "random", "state_seed": 7, "stride": 4099,
                          "measure": {"count": 8, "seed": 42, "ranges": [[19, 19]], "collapse_range": [10, 19]}},
}


def main():
    assert oracle.build_ref(), "compiled reference unavailable"
    for name, spec in SPECS.items():
        with oracle.RefOracle(spec["n"], "sse2") as ref:
            ref.set_multithreading(spec["n"] >= 14)
            got = golden_util.run_case(ref, {"spec": spec})
        np.savez_compressed(os.path.join(HERE, name + ".npz"), spec=np.array(json.dumps(spec)), **got)
        print(name, {k: v.shape for k, v in got.items()})


if __name__ == "__main__":
    main()
