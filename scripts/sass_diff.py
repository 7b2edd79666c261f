"""Compare two builds of libopp_b200.so kernel by kernel (SASS text without the encoding column):
    python scripts/sass_diff.py old.so new.so
Used to prove that a refactor left the already GPU-validated kernels bit-identical."""
import re
import subprocess
import sys


def funcs(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    d, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            d[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip())
    return d


a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
bad = 0
for k in sorted(set(a) | set(b)):
    short = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:90]
    if k not in a:
        print("NEW    ", short)
    elif k not in b:
        print("REMOVED", short)
        bad += 1
    elif a[k] != b[k]:
        print("DIFF   ", short, len(a[k]), len(b[k]))
        bad += 1
print("identical kernels:", sum(1 for k in a if k in b and a[k] == b[k]), "changed/removed:", bad)
sys.exit(1 if bad else 0)
