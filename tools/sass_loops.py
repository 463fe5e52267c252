"""Developer tool: instruction mix of every loop of a kernel (loop = backward-branch target .. branch).
    cuobjdump -sass -fun <mangled kernel> lib.so | python tools/sass_loops.py ["title"]
"""
import re
import sys
from collections import Counter


def op(t):
    t = re.sub(r'^@!?U?P\d+\s+', '', t)
    return t.split()[0]


def mix(body):
    c = Counter(op(t).split('.')[0] for _, t in body)
    full = Counter(op(t) for _, t in body)
    fp64 = c['DFMA'] + c['DMUL'] + c['DADD']
    mov = c['MOV'] + sum(v for k, v in full.items() if k.startswith('IMAD.MOV'))
    return ('%d instructions, FP64 %d (DFMA %d DMUL %d DADD %d), MUFU.RCP64H %d, LDTM %d, STTM %d, LDS %d, STS %d, LDG %d, '
            'STG %d, MOV/IMAD.MOV %d, R2UR %d' % (len(body), fp64, c['DFMA'], c['DMUL'], c['DADD'], c['MUFU'], c['LDTM'], c['STTM'],
                                                   c['LDS'], c['STS'], c['LDG'], c['STG'], mov, c['R2UR']))


def main():
    ins = []
    for l in sys.stdin:
        m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = set()
    for p, t in ins:
        m = re.search(r'BRA\S*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < p:
            loops.add((int(m.group(1), 16), p))
    if len(sys.argv) > 1:
        print(sys.argv[1])
    for a, b in sorted(loops):
        body = [(p, t) for p, t in ins if a <= p <= b]
        if len(body) >= 100:
            print('loop %05x..%05x: %s' % (a, b, mix(body)))
    c = Counter(op(t).split('.')[0] for _, t in ins)
    print('whole kernel: %d instructions; LDTM %d, STTM %d, DFMA %d, MUFU %d, UTMALDG/UBLKCP %d' %
          (len(ins), c['LDTM'], c['STTM'], c['DFMA'], c['MUFU'], c['UTMALDG'] + c['UBLKCP']))


if __name__ == '__main__':
    main()
