'''
Runs every single-kernel / engine parity check of tests/kernel_checks.py on the
GPU, one subprocess per group (so that a faulting kernel cannot poison the CUDA
context of the others), without stopping at the first failure, and writes a
report to gpurun_out/probe.txt.

    python tools/gpu_probe.py [group ...]
'''

import json
import os
import subprocess
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def run_group(name):
    import kernel_checks

    out = []
    for index, check in enumerate(kernel_checks.GROUPS[name]):
        try:
            results = check()
            out.append({'group': name, 'index': index, 'status': 'ok', 'results': results})
        except AssertionError as error:
            out.append({'group': name, 'index': index, 'status': 'MISMATCH', 'detail': str(error)})
        except Exception as error:
            out.append({'group': name, 'index': index, 'status': 'ERROR', 'detail': repr(error),
                        'trace': traceback.format_exc()[-1500:]})
            if 'CUDA' in repr(error) or 'cuda' in repr(error):
                break   # sticky context error: the rest of the group would fail too
    print('@@RESULT@@' + json.dumps(out, default=str))


def main():
    import kernel_checks

    groups = sys.argv[1:] or list(kernel_checks.GROUPS)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    report = []
    for name in groups:
        try:
            proc = subprocess.run([sys.executable, os.path.abspath(__file__), '--group', name], capture_output=True,
                                  text=True, timeout=420)
            lines = [l for l in proc.stdout.splitlines() if l.startswith('@@RESULT@@')]
            if lines:
                report.extend(json.loads(lines[-1][len('@@RESULT@@'):]))
            else:
                report.append({'group': name, 'status': 'CRASH', 'returncode': proc.returncode,
                               'stderr': proc.stderr[-3000:], 'stdout': proc.stdout[-1000:]})
        except subprocess.TimeoutExpired:
            report.append({'group': name, 'status': 'TIMEOUT'})
    failed = 0
    with open(os.path.join(ROOT, 'gpurun_out', 'probe.txt'), 'w') as handle:
        for entry in report:
            if entry['status'] != 'ok':
                failed += 1
            handle.write(json.dumps(entry, default=str) + '\n')
            line = '%-16s #%s %s' % (entry['group'], entry.get('index', '-'), entry['status'])
            if entry['status'] == 'ok':
                line += '  ' + '; '.join('%s rel=%.2e' % (r['name'], r['rel']) for r in entry['results'])
            else:
                line += '  ' + str(entry.get('detail', entry.get('stderr', '')))[:1500]
            print(line)
    print('probe: %d entries, %d not ok' % (len(report), failed))
    return 1 if failed else 0


if __name__ == '__main__':
    if len(sys.argv) == 3 and sys.argv[1] == '--group':
        run_group(sys.argv[2])
    else:
        sys.exit(main())
