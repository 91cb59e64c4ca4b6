"""Summarise a `-Xptxas -v` log: demangled-ish kernel name, registers, spill bytes, smem.  python profiles/ptxas_summary.py <log> [filter]"""
import re, subprocess, sys
log = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
blocks = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", log, re.S)
names = subprocess.run(["c++filt"] + [b[0] for b in blocks], capture_output=True, text=True).stdout.splitlines()
for n, b in zip(names, blocks):
    short = re.sub(r"spk::\(anonymous namespace\)::", "", n)
    short = re.sub(r"\(.*\)$", "", short).replace("void ", "")
    if flt in short:
        print(f"{short:60s} regs {b[4]:>3s} spill st/ld {b[2]:>4s}/{b[3]:>4s} smem {b[5] or 0}")
