'''Dynamic instruction mix of a kernel from an `ncu --set full --import-source on` report: warp instructions executed
per opcode, per score element for the attention kernels.

    python tools/ncu_opmix.py <report.ncu-rep> <kernel substring> [elements]
'''
import collections, csv, re, subprocess, sys

def main():
    path, pattern = sys.argv[1], sys.argv[2]
    elements = float(sys.argv[3]) if len(sys.argv) > 3 else 32 * 16 * 2048 * 2049 / 2 / 32    # warp-elements, bench shape
    raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    byop, total, active = collections.Counter(), 0, False
    for r in rows:
        if r and r[0] == 'Kernel Name':
            active = pattern in r[1]
            hdr = None
            continue
        if not active:
            continue
        if r and r[0] == 'Address':
            hdr = r
            ia, isrc = hdr.index('Instructions Executed'), hdr.index('Source')
            continue
        if hdr is None or len(r) <= ia:
            continue
        try:
            n = int(r[ia])
        except ValueError:
            continue
        m = re.match(r'(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', r[isrc].strip())
        byop[m.group(1) if m else '?'] += n
        total += n
    print('%s: %d warp instructions = %.2f per warp-element' % (pattern, total, total / elements))
    for op, n in byop.most_common(28):
        print('  %-10s %12d  %.2f' % (op, n, n / elements))

main()
