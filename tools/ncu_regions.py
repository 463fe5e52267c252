"""Developer tool: aggregate an ncu source-page CSV (SASS view) into contiguous regions and print, per
region, instructions executed, FP64 share and stall samples.  Regions are split at backward-branch
targets so inner loops show up as their own rows.
    ncu -i x.ncu-rep --page source --csv > x.csv ; python tools/ncu_regions.py x.csv [min_share]
"""
import csv
import re
import sys


def main(path, min_share=0.01):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            addr = int(r[ix['Address']], 16) if r[ix['Address']].startswith('0x') else int(r[ix['Address']])
        except ValueError:
            continue
        data.append((addr, r[ix['Source']], int(r[ix['# Samples']] or 0), int(r[ix['Instructions Executed']] or 0), r))
    base = data[0][0]
    # loop boundaries
    cuts = set()
    for i, (a, src, s, n, r) in enumerate(data):
        m = re.search(r'BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)', src)
        if m:
            tgt = int(m.group(1), 16)
            if tgt >= base:
                tgt -= base          # ncu prints absolute addresses
            if tgt < a - base:
                cuts.add(tgt)
                cuts.add(a - base + 16)
    total_n = sum(d[3] for d in data)
    total_s = sum(d[2] for d in data)
    print('total instructions executed %d, samples %d' % (total_n, total_s))
    regions, cur = [], []
    for d in data:
        if (d[0] - base) in cuts and cur:
            regions.append(cur); cur = []
        cur.append(d)
    if cur:
        regions.append(cur)
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    print('%-14s %6s %9s %6s %6s %6s  top stalls' % ('region', 'instr', 'executed', 'exe%', 'smp%', 'fp64%'))
    for reg in regions:
        n = sum(d[3] for d in reg); s = sum(d[2] for d in reg)
        if n < min_share * total_n and s < min_share * total_s:
            continue
        f = sum(d[3] for d in reg if re.sub(r'^@!?U?P\d+\s+', '', d[1]).split('.')[0].split(' ')[0] in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
        st = {}
        for c in stall_cols:
            v = sum(int(d[4][ix[c]] or 0) for d in reg)
            if v:
                st[c[6:]] = v
        top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
        print('%05x..%05x %7d %9d %6.1f %6.1f %6.1f  %s' % (reg[0][0] - base, reg[-1][0] - base, len(reg), n, 100.0 * n / total_n,
              100.0 * s / max(total_s, 1), 100.0 * f / max(n, 1), ' '.join('%s=%d' % kv for kv in top)))


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.01)
