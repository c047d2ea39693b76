"""Summarise mfem-ad_b200/build/ptxas.log: registers / spills per kernel (optionally filtered by substring)."""
import re, subprocess, sys
pat = sys.argv[1] if len(sys.argv) > 1 else ""
txt = open("mfem-ad_b200/build/ptxas.log").read().split("Compiling entry function '")[1:]
for blk in txt:
    name = blk.split("'")[0]
    try:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except Exception:
        dem = name
    if pat and pat not in dem:
        continue
    regs = re.search(r"Used (\d+) registers", blk)
    spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", blk)
    dem = dem.replace("madb::", "").replace("(int)", "").replace("(unsigned int)", "").replace("(bool)", "")
    print("%4s regs  spill %s/%s  %s" % (regs.group(1) if regs else "?", spill.group(1), spill.group(2), dem[:150]))
