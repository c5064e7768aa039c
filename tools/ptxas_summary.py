"""ptxas_summary.py [NP ...]: registers / stack / spills per kernel from ndt_b200/csrc/build/ptxas_np<NP>.log"""
import re, sys, os
d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ndt_b200", "csrc", "build")
for n in (sys.argv[1:] or ["4", "8", "10"]):
    txt = open(os.path.join(d, f"ptxas_np{n}.log")).read() if n.isdigit() else open(n).read()
    for m in re.finditer(r"Compiling entry function '(\w+)'.*?\n.*?Function properties for \1\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt, re.S):
        name = re.sub(r"^_Z\d+", "", m.group(1))[:28]
        print(f"np{n:>3s} {name:30s} regs {m.group(5):>3s} stack {m.group(2):>5s} spill st/ld {m.group(3)}/{m.group(4)}")
