"""Static audit of the programmatic-dependent-launch rule (csrc/common.cuh): in the SASS of every kernel that executes
griddepcontrol.wait (SASS: ACQBULK), list the global loads (LDG / LD / bulk copies) placed BEFORE the last wait in address order.
Anything listed must be a deliberate pre-wait stream (weights of the GEMV, old K/V rows of the attention kernels) - never
activations produced by the predecessor.  Runs on the build container (no GPU):  python tools/sass_pdl_audit.py [build dir]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bdir = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else os.path.join(ROOT, "chatllm.cpp_b200", "build")
bad = 0
for obj in sorted(f for f in os.listdir(bdir) if f.endswith(".o")):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(bdir, obj)], capture_output=True, text=True).stdout
    for fn in sass.split("Function : ")[1:]:
        name = fn.split("\n", 1)[0].strip()
        ins = re.findall(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", fn)
        waits = [i for i, (_, t) in enumerate(ins) if "ACQBULK" in t]
        if not waits:
            continue
        is_ld = lambda t: re.search(r"\b(LDG|LD\.E|UBLKCP|LDGSTS)", t)
        first = [(a, t.strip()) for a, t in ins[:waits[0]] if is_ld(t)]
        between = [(a, t.strip()) for a, t in ins[waits[0]:waits[-1]] if is_ld(t)]
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
        print(f"{obj}: {demangled}: {len(waits)} wait(s); global loads before the first wait: {len(first)}, between waits: {len(between)}")
        if "-v" in sys.argv:
            for a, t in first + between:
                print(f"      /*{a}*/ {t}")
        bad += len(first) + len(between)
print("total pre-wait loads:", bad)
