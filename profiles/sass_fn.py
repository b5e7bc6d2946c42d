"""Print the SASS of one kernel of a shared library: python profiles/sass_fn.py <lib.so> <substring of the mangled name>"""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, out = None, []
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
        out.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln))
print("\n".join(out))
