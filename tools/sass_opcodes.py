'''
Opcode evidence for the Blackwell claims: per kernel of libcomposer_b200.so, how often the SASS mnemonics that
prove tcgen05 / TMEM / TMA (UTCHMMA, LDTM / STTM, UTMALDG / UTMASTG, UBLKCP), the warp-level tensor path (HMMA) and
the packed-fp32 / MUFU arithmetic occur.  Runs here (cuobjdump needs no GPU):

    python tools/sass_opcodes.py [library] > profiles/sass_opcodes_r2.txt
'''
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'HMMA', 'LDSM', 'LDGSTS', 'MUFU',
         'FFMA2', 'FMUL2', 'FADD2', 'SYNCS', 'REDG', 'RED', 'UCGABAR_ARV', 'STAS', 'MEMBAR', 'CCTL']


def main():
    library = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'composer_b200', 'libcomposer_b200.so')
    sass = subprocess.run(['cuobjdump', '-sass', library], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and name:
            kernels[name][m.group(1)] += 1
            kernels[name]['#'] += 1
    demangled = subprocess.run(['cu++filt'] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    if len(demangled) != len(kernels):
        demangled = list(kernels)
    print('# cuobjdump -sass %s: static instruction counts per kernel (sm_100a)' % os.path.relpath(library, ROOT))
    print('# %-86s %6s  %s' % ('kernel', 'instrs', 'watched opcodes'))
    for (mangled, counts), pretty in sorted(zip(kernels.items(), demangled), key=lambda kv: kv[1]):
        pretty = re.sub(r'\((int|bool|unsigned int)\)', '', pretty)
        pretty = re.sub(r'\(.*', '', pretty).replace('void ', '').replace('cb200::', '')
        seen = ' '.join('%s=%d' % (op, counts[op]) for op in WATCH if counts[op])
        print('%-88s %6d  %s' % (pretty[:88], counts['#'], seen))


if __name__ == '__main__':
    main()
