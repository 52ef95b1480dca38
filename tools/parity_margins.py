'''Prints the measured error of every GPU parity check next to its tolerance (how much margin the tests have).

    python tools/parity_margins.py [group ...]        groups as in tests/kernel_checks.py GROUPS (default: all)
'''
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import kernel_checks as k  # noqa: E402

groups = sys.argv[1:] or list(k.GROUPS)
for group in groups:
    for i, fn in enumerate(k.GROUPS[group]):
        try:
            res = fn()
        except AssertionError as e:
            print('%s-%d FAILED %s' % (group, i, str(e)[:300]))
            continue
        for r in (res if isinstance(res, list) else [res]):
            tol = r.get('tol', 0)
            print('%s-%d  %-72s rel %.3e  tol %.1e  (%.0fx)' % (group, i, r['name'][:72], r['rel'], tol,
                                                                tol / r['rel'] if r['rel'] > 0 else float('inf')))
