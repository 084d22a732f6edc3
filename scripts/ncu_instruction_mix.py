"""Dev tool: executed warp-instructions of the step kernel by phase, from the source page of an ncu capture
(`ncu -i prof.ncu-rep --page source --csv --print-source sass,cuda` > csv).  usage: ncu_instruction_mix.py <csv> <warps>"""
import csv, sys, os
rows = list(csv.reader(open(sys.argv[1])))
warps = float(sys.argv[2]) if len(sys.argv) > 2 else 1024.0
sections, cur = [], None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = {'file': r[1], 'rows': [], 'hdr': None}; sections.append(cur)
    elif cur is not None and r and r[0] == 'Line No':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] and len(r) == len(cur['hdr']):
        cur['rows'].append(r)
seen, per_line = set(), {}
for s in sections:
    if s['file'] in seen:
        break
    seen.add(s['file'])
    h = s['hdr']; li, ie, ai = h.index('Line No'), h.index('Instructions Executed'), h.index('Address')
    name = os.path.basename(s['file'])
    for r in s['rows']:
        if r[ai] == '-' and r[li].isdigit():
            try:
                n = int(float(r[ie] or 0))
            except ValueError:
                n = 0
            per_line[(name, int(r[li]))] = per_line.get((name, int(r[li])), 0) + n
# phases: (file, first line, last line, label) — line numbers of the ROUND-1 sources (the capture behind
# profiles/r01_v8_instruction_mix.txt); re-derive the table from the source page before using it on a newer capture
PH = [
    ('qs_rng.cuh', 1, 10 ** 6, 'Philox blocks + Box-Muller (OU and sensor noise draws)'),
    ('qs_device.cuh', 141, 176, 'state load / store'),
    ('qs_device.cuh', 177, 349, 'dynamics sub-steps (motor lag, thrust / torque, Rodrigues update, integration, floor)'),
    ('qs_device.cuh', 350, 10 ** 6, 'cold device functions (noise re-draw, responses, reset)'),
    ('qs_device.cuh', 1, 140, 'helpers (clamp, norm, shuffles, ballots)'),
    ('qs_step.cuh', 25, 46, 'nearest-pillar distance'),
    ('qs_step.cuh', 47, 193, 'observation rows: self part, K nearest neighbours, 3x3 SDF'),
    ('qs_step.cuh', 194, 224, 'observation tile -> global (coalesced flush)'),
    ('qs_step.cuh', 225, 347, 'reset path, hand-over helpers'),
    ('qs_step.cuh', 348, 514, 'kernel prologue: indices, waits, pillar staging, counters, RNG key'),
    ('qs_step.cuh', 515, 557, 'per-drone step: action -> thrust, 2 sub-steps, reward terms'),
    ('qs_step.cuh', 558, 599, 'all-pairs pass: collisions, proximity, downwash detection'),
    ('qs_step.cuh', 600, 620, 'pillar contact test'),
    ('qs_step.cuh', 621, 693, 'collision / room bookkeeping, counters'),
    ('qs_step.cuh', 694, 715, 'rewards, goal-distance log'),
    ('qs_step.cuh', 716, 794, 'contact responses, scenario tick'),
    ('qs_step.cuh', 795, 850, 'outputs, episode end'),
    ('qs_step.cuh', 851, 10 ** 6, 'observation call site, epilogue'),
]
tot = sum(per_line.values())
acc = [0] * len(PH)
other = 0
for (f, l), n in per_line.items():
    for k, (pf, a, b, _) in enumerate(PH):
        if f == pf and a <= l <= b:
            acc[k] += n
            break
    else:
        other += n
print(f'executed warp-instructions per launch: {tot}  ({tot / warps:.0f} per warp)')
for k in sorted(range(len(PH)), key=lambda k: -acc[k]):
    if acc[k]:
        print(f'{100 * acc[k] / tot:5.1f} %  {acc[k] / warps:7.0f} / warp   {PH[k][3]}')
if other:
    print(f'{100 * other / tot:5.1f} %  {other / warps:7.0f} / warp   other (intrinsics headers, libdevice)')
