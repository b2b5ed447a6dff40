"""SASS evidence for the tensor-core / TMA path, from the built objects (no GPU needed):

    python profiles/sass_listing.py lightglue > profiles/r02_sass_lightglue.txt
    python profiles/sass_listing.py superpoint > profiles/r02_sass_superpoint.txt

Per kernel: instruction count, the Blackwell-specific mnemonics (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UTMALDG = TMA load, UTCBAR = tcgen05.commit, SYNCS = mbarrier, UTCATOMSWS = TMEM alloc) and the first lines carrying them."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
obj = ROOT / "gtsfm_b200" / "csrc" / "_obj" / f"{sys.argv[1]}.o"
out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
KEY = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "LDGSTS", "REDUX", "MUFU", "DFMA", "FFMA2")
fn, per = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        per[fn] = {"n": 0, "hist": collections.Counter(), "lines": []}
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
    if fn and m:
        ins = m.group(1).strip()
        op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0]
        d = per[fn]
        d["n"] += 1
        if op in KEY:
            d["hist"][op] += 1
            if op in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "UTCATOMSWS") and sum(1 for l in d["lines"] if l.split()[0].startswith(op) or op in l) < 2:
                d["lines"].append(ins)
print(f"# cuobjdump -sass {obj.relative_to(ROOT)}  (sm_100a)")
for fn, d in per.items():
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
    if not d["hist"]:
        continue
    print(f"\n{name}: {d['n']} instructions; " + ", ".join(f"{k} x{v}" for k, v in sorted(d["hist"].items())))
    for l in d["lines"]:
        print(f"    {l}")
