"""cuobjdump -sass librstnet_b200.so | python scripts/sass_counts.py profiles/<name>.json
Per-kernel counts of the SASS opcodes that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md: UTCHMMA / UTCQMMA = tcgen05.mma,
UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA tile load / store, LDTM / STTM = tcgen05.ld / st)."""
import collections
import json
import re
import sys

KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "SYNCS", "FFMA", "HMMA", "LDGSTS", "BAR"]
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        cnt[cur][m.group(1).split(".")[0]] += 1
out, tot = {}, collections.Counter()
for f, c in cnt.items():
    sel = {k: c[k] for k in KEYS if c[k]}
    if any(k in sel for k in ("UTCHMMA", "UTCQMMA", "UTMALDG", "LDTM", "STTM", "UTMASTG", "UTCBAR")):
        out[re.sub(r"^_ZN6rstnet\d+", "", f)[:70]] = sel
    for k in KEYS:
        tot[k] += c[k]
out["__total__"] = {k: tot[k] for k in KEYS if tot[k]}
out["__functions__"] = len(cnt)
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out["__total__"]), len(cnt), "functions")
