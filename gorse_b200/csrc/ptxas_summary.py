"""Prints registers / stack / spills per kernel from the build's ptxas -v logs."""
import glob
import os
import re
import subprocess

here = os.path.dirname(os.path.abspath(__file__))
for f in sorted(glob.glob(os.path.join(here, "build", "*.ptxas.log"))):
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
        short = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", short)[:64]
        print(f"{short:64s} regs={m.group(5):>3s} stack={m.group(2):>4s} spill={m.group(3)}/{m.group(4)}")
