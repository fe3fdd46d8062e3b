#!/usr/bin/env python
"""Developer tool: fit the issue-cycle model of the FP64 kernels to an ncu source page.

    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/issue_model.py src.csv [cycles_per_sm]

Model (tools/fp64_microbench.cu measured it on B200): an FP64 instruction holds the sub-partition's
issue port for 2 cycles when it reads at most two distinct 64-bit registers, 3 cycles when it reads
three; every other non-uniform-datapath instruction costs one cycle; U* instructions are free.
"""
import csv
import re
import sys


def parse(src):
    m = re.match(r'\s*(@!?U?P\w+\s+)?(\S+)\s*(.*)', src)
    if not m:
        return None, set()
    op = m.group(2)
    args = m.group(3).rstrip(' ;').split(',')
    rs = set()
    for a in args[1:]:
        a = a.strip().lstrip('-|~!')
        mm = re.match(r'(R\d+)', a)
        if mm:
            rs.add(mm.group(1))
    return op, rs


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    n = dict(fp2=0, fp3=0, uni=0, other=0)
    for r in data:
        ex = int(r[ix['Instructions Executed']])
        if ex == 0:
            continue
        op, rs = parse(r[ix['Source']])
        if op is None:
            continue
        base = op.split('.')[0]
        if base in ('DFMA', 'DMUL', 'DADD', 'DSETP'):
            n['fp3' if len(rs) >= 3 else 'fp2'] += ex
        elif base.startswith('U') or base in ('BRA.U',):
            n['uni'] += ex
        else:
            n['other'] += ex
    smsp = 148 * 4
    cyc = (2 * n['fp2'] + 3 * n['fp3'] + n['other']) / smsp
    print("warp instructions: FP64 with <=2 registers %d, FP64 with 3 registers %d, uniform datapath %d, other %d" % (n['fp2'], n['fp3'], n['uni'], n['other']))
    print("model: %.2f M issue cycles per sub-partition  (FP64 alone at 2 cycles each: %.2f M; 3-register surcharge %.2f M; other %.2f M)"
          % (cyc / 1e6, 2 * (n['fp2'] + n['fp3']) / smsp / 1e6, n['fp3'] / smsp / 1e6, n['other'] / smsp / 1e6))
    if len(sys.argv) > 2:
        print("measured: %.2f M cycles -> model explains %.1f %%" % (float(sys.argv[2]) / 1e6, 100 * cyc / float(sys.argv[2])))


if __name__ == "__main__":
    main()
