#!/usr/bin/env python
"""Per-kernel counts of tensor-core / TMEM / TMA / mbarrier SASS instructions of libvsgb200.so (cuobjdump -sass; no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vidsgg_big_b200", "libvsgb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|UTCATOMSWS|SYNCS\.[A-Z]+)")
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = pat.search(line) if cur else None
    if m:
        counts[cur][m.group(1)] += 1
print("cuobjdump -sass vidsgg_big_b200/libvsgb200.so: tensor-core / TMEM / TMA / mbarrier SASS instructions per kernel (sm_100a)")
print("UTCHMMA = tcgen05.mma (kind::tf32 / kind::f16; .2CTA = cta_group::2), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc,")
print("UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, SYNCS.* = mbarrier operations\n")
tot = collections.Counter()
for k, c in counts.items():
    if not any(op.startswith(("UTC", "LDTM", "UTMA", "UBLK")) for op in c):
        continue
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(CUtensorMap_st.*", "(...)", name)
    name = re.sub(r"\((float const|float \*|const float).*", "(...)", name)[:100]
    print("%-100s %s" % (name, ", ".join("%s x%d" % (op, n) for op, n in sorted(c.items()))))
    tot.update(c)
print("\ntotal: " + ", ".join("%s x%d" % (op, n) for op, n in sorted(tot.items())))
