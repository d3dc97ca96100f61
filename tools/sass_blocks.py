"""Static instruction counts of a kernel's large basic blocks (issue-bound kernels: counts predict time).
usage: python tools/sass_blocks.py OBJ 'kernel name substring' [min_block]"""
import subprocess, sys, re, collections
obj, key = sys.argv[1], sys.argv[2]
minb = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = out.split("Function : ")
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    dem = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip()
    if key not in dem: continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
    targets = set()
    for a, s in ins:
        m = re.search(r"\b(BRA|BRA\.U|BSSY\S*|BRX)\b.*?(0x[0-9a-f]+)", s)
        if m and ("BRA" in s or "BSSY" in s): targets.add(int(m.group(2), 16))
    blocks, cur = [], []
    for a, s in ins:
        if a in targets and cur: blocks.append(cur); cur = []
        cur.append((a, s))
        op = s.split()[1] if s.startswith("@") else s.split()[0]
        if op.startswith(("BRA", "EXIT", "RET", "BRX")): blocks.append(cur); cur = []
    if cur: blocks.append(cur)
    print(dem[:120], "instructions:", len(ins))
    for b in blocks:
        if len(b) < minb: continue
        c = collections.Counter()
        for a, s in b:
            op = s.split()[1] if s.startswith("@") else s.split()[0]
            c[op.split(".")[0]] += 1
        ar = c["FFMA"] + c["FMUL"] + c["FADD"] + c["FFMA2"] + c["FMUL2"] + c["FADD2"]
        print("  block @%04x  n=%d  arith=%d  other=%d  %s" % (b[0][0], len(b), ar, len(b) - ar, dict(c.most_common(12))))
