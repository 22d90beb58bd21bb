"""Static guard for the decode GEMV's inner loop (chatllm.cpp_b200/csrc/gemv.cu), CPU-only: for a FULL row group the RG
rows of a pipeline stage must be one straight-line block, so that their shared-memory loads and dp4a chains interleave.  With a per-row validity
branch nvcc emits one basic block per row (round 2 found exactly that: 4 x 32 dp4a in four blocks, 0.9-1.05 us per stage and warp); this test
fails if a change brings the branches back.  Also: weights and activations reach shared memory only through the bulk-copy engine (UBLKCP)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "chatllm.cpp_b200", "build", "gemv.o")

needs_obj = pytest.mark.skipif(not shutil.which("cuobjdump") or not os.path.exists(OBJ),
                               reason="needs the built gemv.o (python -c 'import __graft_entry__ as g; g.build()') and cuobjdump")


def _kernel(mangled_fragment):
    sass = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    for fn in sass.split("Function : ")[1:]:
        name = fn.split("\n", 1)[0].strip()
        if mangled_fragment in name:
            return name, [t for _, t in re.findall(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", fn)]
    raise AssertionError(f"no kernel matching {mangled_fragment} in gemv.o")


@needs_obj
@pytest.mark.parametrize("fmt,rg,mode,dots_per_row", [("FmtQ4K", 4, 0, 32), ("FmtQ4K", 4, 1, 32), ("FmtQ40", 4, 0, 8), ("FmtQ80", 4, 0, 8)])
def test_full_row_group_is_one_straight_line_block(fmt, rg, mode, dots_per_row):
    name, ins = _kernel(f"gemv_q_kernelINS_6{fmt}ELi{rg}ELi1ELi{mode}E")
    best, run = 0, 0
    for t in ins:
        if re.search(r"\bBRA\b|\bEXIT\b|\bRET\b", t):
            best, run = max(best, run), 0
        elif "IDP.4A" in t:
            run += 1
    best = max(best, run)
    assert best >= rg * dots_per_row, (name, best, "dp4a in the longest branch-free run; expected all rows of a stage in one block")


@needs_obj
def test_shared_memory_is_filled_by_the_bulk_copy_engine_only():
    name, ins = _kernel("gemv_q_kernelINS_6FmtQ4KELi4ELi1ELi0E")
    assert sum("UBLKCP" in t for t in ins) >= 2, name          # weight stages + the activation column
    assert not any(re.search(r"\bLDGSTS\b", t) for t in ins), name
    # no per-thread global -> shared staging loop is left: the only plain global loads are the (rare) bias / expert-id reads of other modes
    assert sum(bool(re.search(r"\bLDG\b", t)) for t in ins) <= 4, (name, [t for t in ins if "LDG" in t])
