'''Prints the logits errors of the generate parity checks next to their tolerances (how much margin the tests have).'''
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import kernel_checks as k  # noqa: E402
for i, fn in enumerate(k.GROUPS['generate']):
    try:
        res = fn()
    except AssertionError as e:
        print(i, 'FAILED', str(e)[:300]); continue
    for r in (res if isinstance(res, list) else [res]):
        if 'logits' in r['name']:
            print(i, r['name'], 'rel %.4f tol %.3f' % (r['rel'], r['tol']))
