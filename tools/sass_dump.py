"""Dump one kernel's SASS (address + instruction) : python tools/sass_dump.py OBJ 'kernel substring' > out"""
import subprocess, sys, re
obj, key = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
for f in out.split("Function : ")[1:]:
    name = f.split("\n", 1)[0]
    dem = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip()
    if key not in dem: continue
    print("//", dem)
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m: print(m.group(1), m.group(2).strip())
