"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump by source line (dev tool)."""
import csv, collections, sys
path = sys.argv[1] if len(sys.argv) > 1 else '/tmp/src.csv'
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rows = list(csv.reader(open(path)))
sections = []; cur = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = {'file': r[1], 'rows': [], 'hdr': None}; sections.append(cur)
    elif cur is not None and r and r[0] == 'Line No':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] and len(r) == len(cur['hdr']):
        cur['rows'].append(r)
seen = set(); use = []
for s in sections:            # first kernel instance only
    if s['file'] in seen: break
    seen.add(s['file']); use.append(s)
tot = collections.Counter(); samp = collections.Counter(); stall = collections.defaultdict(collections.Counter)
for s in use:
    h = s['hdr']; li = h.index('Line No'); ie = h.index('Instructions Executed'); sa = h.index('# Samples'); ai = h.index('Address')
    sc = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    for r in s['rows']:
        if r[ai] == '-' and r[li].isdigit():
            key = (s['file'].split('/')[-1], int(r[li]))
            try:
                tot[key] += int(float(r[ie] or 0)); samp[key] += int(float(r[sa] or 0))
            except ValueError:
                pass
            for i, c in sc:
                try: stall[key][c] += int(float(r[i] or 0))
                except ValueError: pass
T = sum(tot.values()); S = sum(samp.values())
print('total warp-inst', T, 'samples', S)
src = {}
for f in ('qs_device.cuh', 'qs_step.cuh', 'qs_rng.cuh'):
    try: src[f] = open('quad_swarm_rl_b200/csrc/' + f).read().split('\n')
    except OSError: pass
for key, v in samp.most_common(topn):
    f, l = key
    text = src[f][l - 1].strip()[:64] if f in src and l - 1 < len(src[f]) else ''
    top = ', '.join(f'{c[6:]}:{n}' for c, n in stall[key].most_common(3))
    print(f'{f[3:-4]:7s}:{l:4d} samp {v:5d} ({100*v/S:4.1f}%) inst {tot[key]:7d} ({100*tot[key]/T:4.1f}%) | {top} | {text}')
allst = collections.Counter()
for k in stall: allst.update(stall[k])
print('stalls:', ', '.join(f'{c[6:]}:{100*n/S:.1f}%' for c, n in allst.most_common(8)))
pf = collections.Counter(); pfs = collections.Counter()
for (f, l), v in tot.items(): pf[f] += v
for (f, l), v in samp.items(): pfs[f] += v
print({k: f'{100*v/T:.1f}%' for k, v in pf.items()}, {k: f'{100*v/S:.1f}%' for k, v in pfs.items()})
