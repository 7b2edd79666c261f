"""Fingerprint of the kernels in libopp_b200.so: sha256 of each kernel's SASS text.

    python scripts/sass_hash.py write profiles/r1_validated_sass.txt   # after a GPU validation
    python scripts/sass_hash.py check profiles/r1_validated_sass.txt   # later: which kernels changed?

`check` exits non-zero when a kernel listed in the file has different SASS in the current build
(new kernels are reported but allowed) — the way to tell, without a GPU, that an edit left the
GPU-validated kernels bit-identical."""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("OPP_B200_LIB") or os.path.join(ROOT, "onepose_plus_plus_b200", "libopp_b200.so")


def hashes(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    d, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = hashlib.sha256()
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            d[cur].update(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip().encode())
    return {k: v.hexdigest() for k, v in d.items()}


def main():
    mode, path = sys.argv[1], sys.argv[2]
    cur = hashes(LIB)
    if mode == "write":
        with open(path, "w") as f:
            for k in sorted(cur):
                f.write(f"{cur[k]}  {k}\n")
        print(f"wrote {len(cur)} kernel fingerprints to {path}")
        return 0
    ref = dict(reversed(line.split()) for line in open(path) if line.strip())
    bad = [k for k in ref if cur.get(k) != ref[k]]
    new = [k for k in cur if k not in ref]
    for k in bad:
        print("CHANGED" if k in cur else "REMOVED", k)
    for k in new:
        print("new    ", k)
    print(f"{len(ref) - len(bad)} of {len(ref)} fingerprinted kernels unchanged; {len(new)} new")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
