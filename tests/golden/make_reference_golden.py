"""Writes tests/golden/ref_<case>.npz: the outputs of the REFERENCE's own files (/root/reference, unmodified, executed on the
TF1 stand-in tests/golden/tf1_shim.py) for the cases of tests/golden/reference_cases.py.  Run here, where the reference tree
exists; the fixtures travel with the repo and tests/test_oracle_vs_reference_cpu.py holds the oracle to them.

    python tests/golden/make_reference_golden.py [case ...]
"""
import sys
import time

import numpy as np

import reference_cases as rc


def write_json_fixtures():
    """models/config.py defaults and the signatures of every function the reference's path modules define."""
    import json
    import os
    import reference_runner as rr
    for fname, obj in (('ref_config_defaults.json', rr.reference_config_defaults()), ('ref_signatures.json', rr.reference_signatures()),
                       ('ref_variables.json', rr.reference_variables()), ('ref_restore_maps.json', rr.reference_restore_maps())):
        with open(os.path.join(rc.HERE, fname), 'w') as f:
            json.dump(obj, f, indent=0, sort_keys=True)
        print('wrote', fname)


def main(argv):
    if not argv:
        write_json_fixtures()
    names = argv or sorted(rc.CASES)
    for name in names:
        t0 = time.time()
        ref = rc.CASES[name][0]()
        np.savez_compressed(rc.fixture_path(name), **ref)
        bad = rc.compare(rc.CASES[name][1](), ref)
        print('%-34s %4d entries, %.1f s, oracle vs reference: %s' % (name, len(ref), time.time() - t0, 'identical' if not bad else bad[:6]))


if __name__ == '__main__':
    main(sys.argv[1:])
