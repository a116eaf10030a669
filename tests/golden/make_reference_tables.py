#!/usr/bin/env python
"""tests/golden/reference_tables.npz = the shape tables integration/gpu_backend.py::extract_tables reads out of the
REFERENCE'S OWN objects (closures of polygon.py / ball.py / square.py) for the main.py map, frozen so that the GPU box
(no /root/reference) can run the raw ctypes binding of INTEGRATION.md section 2 against the reference's golden costs.

Run only in the authoring container:   python tests/golden/make_reference_tables.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, 'integration'))


def reference_map_and_problem(N=80):
    """the reference's RegionMap + Problem for the main.py scenario, built by the reference's own constructors"""
    import make_golden_integral as mgi
    spec, problem = mgi._setup()
    pr = problem(N)
    return spec, pr.map, pr


def main():
    import gpu_backend
    spec, m, pr = reference_map_and_problem()
    t = gpu_backend.extract_tables(m)
    # a square() too, so that the box-side probe is frozen as well (the main.py map has none)
    from square import square
    sq = square([1.0, 1.0], 0.5, 0.25)
    t['square_records'] = np.array([gpu_backend.inequality_record(f) for f in sq.inequalities])
    np.savez_compressed(os.path.join(HERE, 'reference_tables.npz'), **t)
    print({k: np.asarray(v).shape for k, v in t.items()})


if __name__ == '__main__':
    main()
