"""Reads the stderr of a run with B200_WAVE_TRACE=1 (Ldu::waveSweep) and prints, for a few tiles of sweep number N, the globaltimer
stamps (ns) and the clock cycles each warp spent per phase.  Usage: python tools/wave_trace.py FILE [sweep index]"""
import sys

import numpy as np

fn = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 5
blocks, cur = [], None
for l in open(fn):
    if l.startswith('WAVE_TRACE'):
        p = l.split(); cur = {'op': int(p[2]), 'nJ': int(p[4]), 'nK': int(p[6]), 'rows': {}}; blocks.append(cur)
    elif l.startswith('WT'):
        p = l.split(); cur['rows'][(int(p[1]), int(p[2]))] = [int(x) for x in p[3:]]
print(len(blocks), 'sweeps:', [b['op'] for b in blocks[:10]])
for b in blocks[which:which + 2]:
    r, op, nJ, nK = b['rows'], b['op'], b['nJ'], b['nK']
    print('op', op)
    first = (0, 0) if op != 2 else (nJ - 1, nK - 1)
    sgn = 1 if op != 2 else -1
    keys = [first] + [(first[0] + sgn * d, first[1]) for d in (1, 2, 10) if 0 <= first[0] + sgn * d < nJ]
    keys += [(nJ // 2, nK // 2), (nJ - 1 - first[0], nK - 1 - first[1])]
    for key in dict.fromkeys(keys):
        v = r[key]
        print(key, 'start', v[11], 'ready', v[0:5], 'done', v[5:10], 'last', v[10], 'dur', v[10] - v[11])
        print('    mem: waitEmpty %d bulk %d loads %d poll %d smem+arrive %d side %d  pollpasses %d | compute: waitFull %d work %d' % tuple(v[12:19] + v[20:22]))
    if nJ > 1:
        d = [r[(J, first[1])][10] for J in (range(nJ) if op != 2 else range(nJ - 1, -1, -1))]
        print('J-lag of last group: mean %.0f ns' % np.mean(np.diff(d)), np.diff(d)[:10])
    if nK > 1:
        d = [r[(first[0], K)][10] for K in (range(nK) if op != 2 else range(nK - 1, -1, -1))]
        print('K-lag of last group: mean %.0f ns' % np.mean(np.diff(d)), np.diff(d)[:10])
    print('sweep end', max(v[10] for v in r.values()))
