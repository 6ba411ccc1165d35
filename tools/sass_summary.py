"""Writes profiles/r2_sass_tensor_tma_mnemonics.txt: per kernel of libdpc_b200.so, the counts of the tensor-core / TMEM / TMA /
mbarrier SASS mnemonics (cuobjdump -sass; no GPU needed) — the evidence that the hot path is tcgen05 / TMEM / TMA code."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "diffphycon_b200", "libdpc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r'\b(UTCHMMA[.\w]*|UTCQMMA[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTCBAR[.\w]*|UTCCP[.\w]*|HMMA[.\w]*|'
                 r'UTMAPF[.\w]*|SYNCS[.\w]*|DFMA|DADD|DMUL)\b')
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        m = pat.search(line)
        if m:
            counts[cur][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
lines = ["# cuobjdump -sass diffphycon_b200/libdpc_b200.so — per kernel, counts of the tensor-core / TMEM / TMA / mbarrier mnemonics",
         "# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2); LDTM / STTM = tcgen05.ld / st; UTMALDG / UTMASTG = cp.async.bulk.tensor load /",
         "# store; UTCBAR = tcgen05.commit; SYNCS = mbarrier ops; HMMA = legacy mma.sync; DFMA/DADD/DMUL = fp64 (rollout CG)", ""]
for (f, c), name in zip(counts.items(), names):
    if not c:
        continue
    name = re.sub(r'\(.*', '', name)
    lines.append(name)
    lines.append("    " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
open(os.path.join(ROOT, "profiles", "r2_sass_tensor_tma_mnemonics.txt"), "w").write("\n".join(lines) + "\n")
print(len(lines) // 2, "kernels")
