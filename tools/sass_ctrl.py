"""Developer tool: decode the scheduling control word of each SASS instruction of a kernel's hot loops.

    python tools/sass_ctrl.py <object> <kernel-name-substring> [-v]

For every loop with >= 20 FP64 instructions: the sum of the ptxas-assigned issue stalls (the minimum
cycles a lone warp needs per trip before any scoreboard wait), and with -v one line per instruction
(stall, yield, write/read barrier, wait mask, reuse flags).  Volta+ encoding: bits 105..125 of the
128-bit instruction word.
"""
import re
import subprocess
import sys

FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP')


def parse(path, kname):
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    cur, body = None, []
    lines = out.splitlines()
    i = 0
    while i < len(lines):
        line = lines[i]
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
        elif cur and kname in cur:
            m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/', line)
            if m and i + 1 < len(lines):
                m2 = re.search(r'/\* (0x[0-9a-f]+) \*/', lines[i + 1])
                hi = int(m2.group(1), 16) if m2 else 0
                body.append((int(m.group(1), 16), m.group(2).strip(), hi))
                i += 1
        i += 1
    return body


def ctrl(hi):
    return dict(stall=(hi >> 41) & 0xf, yld=(hi >> 45) & 1, wbar=(hi >> 46) & 7, rbar=(hi >> 49) & 7,
                wait=(hi >> 52) & 0x3f, reuse=(hi >> 58) & 0xf)


def main():
    path, kname = sys.argv[1], sys.argv[2]
    verbose = '-v' in sys.argv
    body = parse(path, kname)
    idx = {a: i for i, (a, _, _) in enumerate(body)}
    for i, (a, text, hi) in enumerate(body):
        m = re.search(r'BRA\s+(?:U?!?U?P\d+,\s*)?`?\(?\.?L?_?x?_?\d*\)?|BRA', text)
        m = re.search(r'0x([0-9a-f]+)', text) if 'BRA' in text else None
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a or tgt not in idx:
            continue
        s = idx[tgt]
        instrs = body[s:i + 1]
        nf = sum(1 for _, t, _ in instrs if re.sub(r'^@!?U?P\d+\s+', '', t).split('.')[0].split()[0] in FP64)
        if nf < 20:
            continue
        stalls = sum(ctrl(h)['stall'] for _, _, h in instrs)
        waits = sum(1 for _, _, h in instrs if ctrl(h)['wait'])
        print('loop @%05x..%05x: %d instr, %d FP64, stall sum %d cycles (FP64 pipe %.0f%% if nothing else waits), '
              '%d instrs wait on a barrier' % (tgt, a, len(instrs), nf, stalls, 200.0 * nf / max(stalls, 1), waits))
        if verbose:
            for ad, t, h in instrs:
                c = ctrl(h)
                print('  %05x  s%-2d %s w%s r%s wait=%02x reuse=%x  %s' % (
                    ad, c['stall'], 'Y' if c['yld'] else '-', c['wbar'] if c['wbar'] != 7 else '-',
                    c['rbar'] if c['rbar'] != 7 else '-', c['wait'], c['reuse'], t))


if __name__ == '__main__':
    main()
