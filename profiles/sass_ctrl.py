"""SASS of one kernel with decoded control fields (stall, yield, write/read barrier, wait mask):
python profiles/sass_ctrl.py <lib.so> <substring of mangled name> [first_idx last_idx]"""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi_ = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, idx, pend = None, -1, None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); idx = -1; continue
    if not (cur and pat in cur):
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*/\* (0x[0-9a-f]+) \*/", ln)
    if m:
        idx += 1
        pend = (idx, m.group(1), m.group(2))
        continue
    m = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", ln)
    if m and pend:
        h = int(m.group(1), 16)
        stall, yld, wbar, rbar, wait = (h >> 41) & 0xf, (h >> 45) & 1, (h >> 46) & 7, (h >> 49) & 7, (h >> 52) & 0x3f
        i, addr, ins = pend
        if lo <= i <= hi_:
            w = "".join(str(b) if (wait >> b) & 1 else "-" for b in range(6))
            print(f"{i:5d} {addr} st={stall:2d} y={yld} W={wbar if wbar != 7 else '-'} R={rbar if rbar != 7 else '-'} wait={w}  {ins}")
        pend = None
