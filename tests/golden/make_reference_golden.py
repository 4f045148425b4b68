"""Writes tests/golden/ref_<case>.npz: the outputs of the REFERENCE's own files (/root/reference, unmodified, executed on the
TF1 stand-in tests/golden/tf1_shim.py) for the cases of tests/golden/reference_cases.py.  Run here, where the reference tree
exists; the fixtures travel with the repo and tests/test_oracle_vs_reference_cpu.py holds the oracle to them.

    python tests/golden/make_reference_golden.py [case ...]
"""
import sys
import time

import numpy as np

import reference_cases as rc


def main(argv):
    names = argv or sorted(rc.CASES)
    for name in names:
        t0 = time.time()
        ref = rc.CASES[name][0]()
        np.savez_compressed(rc.fixture_path(name), **ref)
        bad = rc.compare(rc.CASES[name][1](), ref)
        print('%-34s %4d entries, %.1f s, oracle vs reference: %s' % (name, len(ref), time.time() - t0, 'identical' if not bad else bad[:6]))


if __name__ == '__main__':
    main(sys.argv[1:])
