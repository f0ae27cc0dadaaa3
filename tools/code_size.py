#!/usr/bin/env python
"""Static code size of one kernel instantiation by source function (SASS instructions, 16 B each):
where the bytes are that compete for the 32 KB L1.5 instruction cache.
    python tools/code_size.py [mangled-substring, default k_wavefront2ILi512ELi0E]"""
import sys
from collections import Counter

import ncu_lines

kernel = sys.argv[1] if len(sys.argv) > 1 else "k_wavefront2ILi512ELi0E"
table = ncu_lines.line_table(kernel)
funcs = ncu_lines.function_of_line()
by_func = Counter()
for addr, key in table.items():
    name = None
    if key:
        t = funcs.get(key[0])
        if t and key[1] - 1 < len(t):
            name = t[key[1] - 1]
    by_func[(key[0] if key else None, name)] += 1
total = sum(by_func.values())
print("%s: %d instructions = %.1f KB" % (kernel, total, total * 16 / 1024.0))
for (f, n), c in by_func.most_common(45):
    print("  %-18s %-28s %5d  %5.1f KB" % (f, n, c, c * 16 / 1024.0))
