"""Per-kernel hash of the SASS of a built libgpifdtd.so (cuobjdump -sass): `python scripts/sass_hashes.py lib.so > hashes.txt`.
Used to show that a source change leaves the machine code of kernels it is not meant to touch identical."""
import hashlib
import re
import subprocess
import sys

out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True, check=True).stdout
cur, body, res = None, [], {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if cur:
            res[cur] = body
        cur, body = m.group(1), []
    elif cur:
        # drop the hex encodings (/* 0x... */) and address columns, keep the mnemonics and operands
        t = re.sub(r"/\*\s*[0-9a-fx]+\s*\*/", "", line).strip()
        t = re.sub(r"c\[0x0\]\[0x[0-9a-f]+\]", "c[0x0][.]", t)      # kernel-parameter offsets move when a parameter struct grows
        if t:
            body.append(t)
if cur:
    res[cur] = body
for k in sorted(res):
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    print(hashlib.sha1("\n".join(res[k]).encode()).hexdigest()[:16], len(res[k]), name)
