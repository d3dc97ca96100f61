"""Opcode mix of a kernel from an ncu source page: python tools/sass_mix.py REP KERNEL_REGEX [min_exec_fraction]
prints executed warp instructions per opcode (hot part: instructions executed at least FRACTION of the maximum count)."""
import csv, io, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
iSm = hdr.index("# Samples")
mix = collections.Counter(); tot = 0; samp = collections.Counter()
ex = [int(r[iE]) for r in rows[2:] if len(r) > iE and r[iE].isdigit()]
mx = max(ex)
for r in rows[2:]:
    if len(r) <= iE or not r[iE].isdigit(): continue
    e = int(r[iE])
    if e < frac * mx: continue
    src = r[iS].strip()
    if src.startswith("@"): src = src.split(None, 1)[1]
    op = src.split()[0].split(".")[0]
    mix[op] += e; tot += e; samp[op] += int(r[iSm] or 0)
print("total executed", tot, "max per-instruction", mx)
for op, e in mix.most_common(30):
    print("%-10s %14d  %6.2f %%  samples %d" % (op, e, 100.0 * e / tot, samp[op]))
