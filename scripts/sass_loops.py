#!/usr/bin/env python3
"""Per-loop opcode statistics of one kernel from `cuobjdump -sass` output (compile-time check before spending GPU time).
usage: cuobjdump -sass lib.so | python scripts/sass_loops.py k_trimersILi1"""
import re
import sys

txt = sys.stdin.read()
want = sys.argv[1]
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    if want not in f.split('\n')[0]:
        continue
    ops = []
    for l in f.split('\n'):
        m = re.search(r'/\*([0-9a-f]{4,6})\*/\s+(.*?);', l)
        if m:
            ops.append((int(m.group(1), 16), m.group(2)))
    tot = lambda p, b: sum(1 for _, o in b if re.search(p, o))
    print("function: %d instr, DFMA %d DMUL %d DADD %d LDS %d STS %d LDL %d STL %d" % (
        len(ops), tot(r'\bDFMA', ops), tot(r'\bDMUL', ops), tot(r'\bDADD', ops), tot(r'\bLDS', ops), tot(r'\bSTS', ops), tot(r'\bLDL', ops), tot(r'\bSTL', ops)))
    for a, o in ops:
        m = re.search(r'BRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)', o)
        if m and int(m.group(1), 16) < a:
            t = int(m.group(1), 16)
            body = [(x, y) for x, y in ops if t <= x <= a]
            print("loop %6x-%6x n=%4d DFMA=%3d DMUL=%2d LDS=%3d STS=%2d LDL=%2d STL=%2d pred-mem=%d other=%d" % (
                t, a, len(body), tot(r'\bDFMA', body), tot(r'\bDMUL', body), tot(r'\bLDS', body), tot(r'\bSTS', body), tot(r'\bLDL', body),
                tot(r'\bSTL', body), tot(r'@!?P\d\s+(LDS|STS)', body),
                len(body) - tot(r'\b(DFMA|DMUL|DADD|LDS|STS)\b', body)))
