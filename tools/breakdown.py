#!/usr/bin/env python
"""Per-class summary of a bench.py --breakdown JSON."""
import json, collections, sys
rows = json.load(open(sys.argv[1]))
tot = sum(r['ms'] for r in rows)
cls = collections.OrderedDict()
for r in rows:
    k = r['kernel'].replace(' past', '').replace(' fut', '')
    c = cls.setdefault(k, [0, 0, 0]); c[0] += r['ms']; c[1] += r['alg_MB']; c[2] += 1
print('total %.3f ms' % tot)
for k, (ms, mb, n) in sorted(cls.items(), key=lambda kv: -kv[1][0]):
    print('%-26s n=%d %7.4f ms %5.1f%%  alg %7.1f MB %7.1f GB/s  floor@6536GB/s %.4f ms  frac %.2f' % (
        k, n, ms, 100 * ms / tot, mb, mb / ms, mb / 6536, mb / 6536 / ms))
