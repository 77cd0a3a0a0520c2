"""Condense the ptxas -v logs of the last build (casmvsnet_pl_b200/csrc/build/*.ptxas.log,
untracked) into one tracked table: registers / spills / static smem per kernel instantiation.

    python profiles/ptxas_summary.py > profiles/ptxas_summary.txt
"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for log in sorted(glob.glob(os.path.join(ROOT, "casmvsnet_pl_b200", "csrc", "build", "*.ptxas.log"))):
    t = open(log).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r"ptxas info\s+: Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", t):
        name = m.group(1)
        try:
            name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        except Exception:
            pass
        name = re.sub(r"\(.*$", "", name).replace("casmvs::", "")
        rows.append((os.path.basename(log).replace(".ptxas.log", ""), name, int(m.group(5)),
                     int(m.group(3)), int(m.group(4)), m.group(8) or "0"))
print(f"{'file':<18}{'regs':>5}{'spill_st':>9}{'spill_ld':>9}{'smem':>8}  kernel")
for f, n, r, ss, sl, sm in rows:
    print(f"{f:<18}{r:>5}{ss:>9}{sl:>9}{sm:>8}  {n[:150]}")
print(f"# {len(rows)} kernel instantiations; spilling ones: {sum(1 for r in rows if r[3] or r[4])}")
