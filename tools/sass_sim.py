"""Developer tool: single-warp in-order issue model of a SASS loop body (no GPU needed).

Finds the innermost hot loop(s) of a kernel in `cuobjdump -sass` output (a backward branch and its
target), then walks the body with a scoreboard: an instruction issues when its source registers
are ready and its pipe is free.  Latencies measured on B200 (tools/latency_probe.py): DFMA/DMUL/
DADD 8.7 cycles dependent issue and one per 2 cycles per SM sub-partition; MUFU.RCP64H 17.5;
LDS ~30 (B300_MICROARCH.md).  Prints cycles per trip and the FP64-pipe utilisation a lone warp
would reach -- the quantity the EM kernel needs to maximise.

    python tools/sass_sim.py <object-or-so> <kernel-name-substring>
"""
import re
import subprocess
import sys

LAT = {'DFMA': 9, 'DMUL': 9, 'DADD': 9, 'DSETP': 9, 'MUFU': 18, 'LDS': 30, 'LDG': 400, 'LDC': 30, 'LDCU': 30,
       'SHFL': 26, 'S2R': 20}
FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP')


def parse(path, kname):
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    cur, body = None, []
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            continue
        if cur and kname in cur:
            m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', line)
            if m:
                body.append((int(m.group(1), 16), m.group(2).strip()))
    return body


def regs_of(tok):
    out = []
    for m in re.finditer(r'\b(U?R)(\d+)(\.64)?\b', tok):
        out.append((m.group(1), int(m.group(2))))
    return out


def simulate(instrs):
    ready = {}
    t = 0
    fp64_free = 0
    n_fp64 = 0
    for text in instrs:
        pred = None
        txt = text
        m = re.match(r'(@!?U?P\d+)\s+(.*)', txt)
        if m:
            txt = m.group(2)
        op = txt.split()[0]
        base = op.split('.')[0]
        operands = txt[len(op):].split(',')
        wide = base in FP64 or base in ('LDS', 'STS', 'MUFU') and '.64' in op or base in ('DFMA', 'DMUL', 'DADD')
        dst = regs_of(operands[0]) if operands and base not in ('STS', 'STG', 'BRA', 'BSYNC', 'BSSY', 'ISETP') else []
        srcs = []
        for o in (operands[1:] if dst else operands):
            srcs += regs_of(o)
        def expand(rs, w):
            ex = []
            for kind, n in rs:
                ex.append((kind, n))
                if w and kind == 'R':
                    ex.append((kind, n + 1))
            return ex
        is64 = base in ('DFMA', 'DMUL', 'DADD', 'DSETP') or '.64' in op
        srcs = expand(srcs, is64)
        dsts = expand(dst, is64 and base != 'MUFU')
        start = t + 1
        for r in srcs:
            start = max(start, ready.get(r, 0))
        if base in FP64:
            start = max(start, fp64_free)
            fp64_free = start + 2
            n_fp64 += 1
        t = start
        lat = LAT.get(base, 5)
        for r in dsts:
            ready[r] = t + lat
    return t, n_fp64


def main():
    path, kname = sys.argv[1], sys.argv[2]
    body = parse(path, kname)
    addr_index = {a: i for i, (a, _) in enumerate(body)}
    loops = []
    for i, (a, text) in enumerate(body):
        m = re.search(r'BRA\s+(?:U?P?\d*,?\s*)?0x([0-9a-f]+)', text)
        if m and 'BRA' in text:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    print('%d instructions, %d backward branches' % (len(body), len(loops)))
    for s, e in loops:
        instrs = [t for _, t in body[s:e + 1]]
        n = len(instrs)
        nf = sum(1 for t in instrs if re.sub(r'^@!?U?P\d+\s+', '', t).split('.')[0].split()[0] in FP64)
        if nf < 20:
            continue
        # steady state: simulate two trips and take the second
        t1, _ = simulate(instrs)
        t2, f2 = simulate(instrs + instrs)
        trip = t2 - t1
        print('loop @%05x..%05x: %4d instr, %3d FP64, %5d cycles/trip alone -> FP64 pipe %.0f%%, issue %.0f%%' %
              (body[s][0], body[e][0], n, nf, trip, 100.0 * 2 * nf / trip, 100.0 * n / trip))


if __name__ == '__main__':
    main()
