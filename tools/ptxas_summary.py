"""Prints registers / spills per kernel from gbwt-rs_b200/build_ptxas.log (optionally only names containing argv[1:])."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = open(os.path.join(ROOT, "gbwt-rs_b200", "build_ptxas.log")).read()
for b in re.split(r"ptxas info\s+: Compiling entry function '", log)[1:]:
    name = b.split("'")[0]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(gbwt_b200::IndexView.*", "", dem).replace("gbwt_b200::", "").replace("(anonymous namespace)::", "")
    if sys.argv[1:] and not any(a in dem for a in sys.argv[1:]):
        continue
    used = re.search(r"Used (\d+) registers", b).group(1)
    sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b).groups()
    print(f"{dem:70s} regs {used:>3s}  stack {sp[0]:>3s}  spill st/ld {sp[1]}/{sp[2]}")
