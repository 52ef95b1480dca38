'''Aggregates an `ncu --page source --csv` dump by SASS opcode: executed warp instructions and stall samples.'''
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, iex, ismp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops = collections.Counter(); smp = collections.Counter()
for r in rows[2:]:
    if len(r) <= iex or not r[iex].isdigit(): continue
    text = r[isrc].strip()
    if text.startswith('@'): text = text.split(None, 1)[1]
    op = text.split()[0].split('.')[0] if text else '?'
    full = text.split()[0]
    key = op if op not in ('MUFU', 'LDSM', 'MOVM') else full
    ops[key] += int(r[iex] or 0); smp[key] += int(r[ismp] or 0)
total = sum(ops.values()); stotal = sum(smp.values())
print('total warp instructions %d, samples %d' % (total, stotal))
for k, v in ops.most_common(28):
    print('%-22s %12d %5.1f%%   stall samples %5.1f%%' % (k, v, 100.0 * v / total, 100.0 * smp[k] / max(stotal, 1)))
