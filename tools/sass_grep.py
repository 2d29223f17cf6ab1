"""Counts of the SASS mnemonics that prove (or rule out) the Blackwell-native paths, per kernel of librio_cuda.so
(B200_PROFILING.md "What proves a Blackwell-native kernel").  usage: python tools/sass_grep.py > profiles/r02_sass_grep.txt"""
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "rio_rs_b200/librio_cuda.so"
WANT = ["UTCHMMA", "LDTM", "UTCBAR", "SYNCS", "UBLKCP", "UTMALDG", "VIMNMX3", "IMAD", "HMMA", "LDS", "LDG", "ATOMS", "REDG"]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
rows, name, cnt = [], None, None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        if name:
            rows.append((name, cnt))
        name, cnt = m.group(1), dict.fromkeys(WANT, 0)
        continue
    if name:
        mm = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if mm:
            op = mm.group(1)
            for w in WANT:
                if op == w or op.startswith(w):
                    cnt[w] += 1
if name:
    rows.append((name, cnt))
dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass %s: per kernel, instructions whose opcode is UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier),"
      " UBLKCP (cp.async.bulk = TMA bulk copy), UTMALDG (TMA tensor load), VIMNMX3 (3-input integer min/max), IMAD, HMMA (legacy mma.sync: must be 0), LDS, LDG, ATOMS, REDG" % so)
print("kernel," + ",".join(WANT))
for (raw, cnt), d in zip(rows, dem):
    d = d.replace("(anonymous namespace)::", "").replace("rio::", "").replace("void ", "")
    d = re.sub(r"\(.*$", "", d).replace(",", ";")
    print(d + "," + ",".join(str(cnt[w]) for w in WANT))
