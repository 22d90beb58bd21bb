"""Static guard for the programmatic-dependent-launch rule (chatllm.cpp_b200/csrc/common.cuh), CPU-only: in the SASS
of every kernel that executes griddepcontrol.wait (ACQBULK) no global load may precede the wait, except the deliberate pre-wait
streams — the GEMV's bulk weight copies and the old K / V rows of the two split attention kernels.  A `const T * __restrict__`
parameter on producer-written data lets nvcc hoist loads above the wait (it did, round 1); this test catches the next one at build time."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "chatllm.cpp_b200", "build")
OBJS = ["fused.o", "gemv.o", "ops.o", "ops_generic.o", "quantize.o", "prefill.o", "prefill_tc.o", "decode_mk.o"]

needs_objs = pytest.mark.skipif(not shutil.which("cuobjdump") or not all(os.path.exists(os.path.join(BUILD, o)) for o in OBJS),
                                reason="needs the built kernel objects (python -c 'import __graft_entry__ as g; g.build()') and cuobjdump")

LOAD = re.compile(r"\b(LDG|LD\.E|UBLKCP|LDGSTS)")
STORE = re.compile(r"\b(STG|ST\.E|RED\.E|ATOMG)")


def kernels(obj):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True, check=True).stdout
    for fn in sass.split("Function : ")[1:]:
        name = fn.split("\n", 1)[0].strip()
        ins = [t for _, t in re.findall(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", fn)]
        yield name, ins


@needs_objs
@pytest.mark.parametrize("obj", OBJS)
def test_no_producer_load_is_hoisted_above_the_pdl_wait(obj):
    checked = 0
    for name, ins in kernels(obj):
        waits = [i for i, t in enumerate(ins) if "ACQBULK" in t]
        if "decode_mk_kernel" in name:
            # the persistent kernel is a cooperative launch WITHOUT the programmatic-serialization attribute: plain stream order
            assert not waits, name
            continue
        # every other kernel is launched through launch_pdl: one that never waits could finish before its predecessor and break the
        # transitive ordering of the chain (repack_bytes_kernel had no wait and the old audit skipped it silently)
        assert waits, f"{name}: launched with programmatic stream serialization but never executes griddepcontrol.wait"
        checked += 1
        stores = [t for t in ins[:waits[0]] if STORE.search(t)]
        assert not stores, (name, "global stores before the first wait", stores)
        before = [t for t in ins[:waits[0]] if LOAD.search(t)]
        between = [t for t in ins[waits[0]:waits[-1]] if LOAD.search(t)]
        if "gemv_q_kernel" in name:
            # weights only, and only through the bulk-copy engine; activations (plain LDG) come after the wait
            assert all("UBLKCP" in t for t in before) and not between, (name, before, between)
        elif "attn_pv_split_kernel" in name or "attn_pv_cluster_kernel" in name:
            # one round of V fragments (2 k-groups x 4 chunks x 2 rows) for the splits that do not hold the newest position
            assert len(before) <= 16 and all("LDG.E.128" in t for t in before) and not between, (name, before, between)
        elif "attn_scores_mma_kernel" in name:
            # K fragments between the two (mutually exclusive) waits; q strictly after
            assert not before and len(between) <= 8 and all("LDG.E.128" in t for t in between), (name, before, between)
        else:
            assert not before and not between, (name, before, between)
    if obj in ("fused.o", "gemv.o"):
        assert checked > 0


def test_tcgen05_smem_layout_matches_the_cute_canonical_layout(tmp_path):
    """csrc/prefill_tc.cu writes its operand tiles with tc_off() and encodes LBO = 2048 / SBO = 128 by hand.  tools/cute_layout_check.cu
    (host code, compiled against the CUTLASS/CuTe headers vendored in the image) checks both against CuTe's own canonical UMMA K-major
    no-swizzle layout and against what cute::UMMA::make_umma_desc would derive for it."""
    import glob
    inc = glob.glob("/opt/prime-rl/.venv/lib/python*/site-packages/flashinfer/data/cutlass/include")
    if not inc or not shutil.which("nvcc"):
        pytest.skip("needs nvcc and the vendored CUTLASS headers")
    exe = str(tmp_path / "layout_check")
    subprocess.run(["nvcc", "-std=c++17", "-I" + inc[0], "-o", exe, os.path.join(ROOT, "tools", "cute_layout_check.cu")], check=True, capture_output=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "mismatches: 0" in out
    assert "SBO (uint128 units) = 8 " in out and "LBO (uint128 units) = 128 " in out
    assert "offset of (row 0, k 32) = 4096 bytes" in out
    assert out.count("EQUAL") == 4 and "DIFFERENT" not in out, out   # smem + instruction descriptor bit patterns, N = 128 and N = 64 tiles
    assert "tc_off64 vs tiled canonical layout mismatches: 0" in out and "LBO (uint128 units) = 64 " in out
    tiled = [l for l in out.splitlines() if l.startswith("tiled")][0].split(" o ")[-1]
    mine = [l for l in out.splitlines() if l.startswith("mine")][0].split(": ")[-1]
    assert tiled.strip() == mine.strip(), (tiled, mine)
