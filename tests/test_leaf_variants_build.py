"""The leaf kernel's tile fill: the default build uses TMA 1-D bulk copies with mbarrier completion (UBLKCP / SYNCS.PHASECHK in the
SASS; measured in round 2, profiles/r02a_call.log), -DNBODY_LEAF_BULK=0 builds the per-lane cp.async rows (LDGSTS) it replaced for
A/B runs. Both must keep compiling for sm_100a within the leaf kernel's register budget and without spills."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "nbody_b200", "csrc", "leaf.cu")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def compile_leaf(tmp_path, tag, defs):
    obj = str(tmp_path / f"leaf_{tag}.o")
    r = subprocess.run([NVCC, "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--expt-relaxed-constexpr",
                        "-Xptxas", "-v"] + defs + ["-c", SRC, "-o", obj], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    log = r.stdout + r.stderr
    res = {}
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads"
                         r".*?Used (\d+) registers", log, re.S):
        res[m.group(1)] = (int(m.group(2)), int(m.group(3)) + int(m.group(4)), int(m.group(5)))
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    per_kernel, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per_kernel[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            per_kernel[name].append(line)
    return res, per_kernel


def count(lines, pattern):
    return sum(1 for l in lines if re.search(pattern, l))


@pytest.mark.parametrize("tag,defs", [("default", []), ("rowfill", ["-DNBODY_LEAF_BULK=0"]),
                                      ("bulk_rows2", ["-DNBODY_LEAF_ROWS=2"])])
def test_variant_builds_within_budget_and_contains_its_instructions(tmp_path, tag, defs):
    res, sass = compile_leaf(tmp_path, tag, defs)
    leaf = {k: v for k, v in res.items() if "6k_leafILi" in k}
    assert len(leaf) == 8                                     # orders 2, 3, 4, 5 x softened / unsoftened
    for name, (stack, spill, regs) in leaf.items():
        assert regs <= 128 and spill == 0 and stack == 0, (tag, name, stack, spill, regs)
    for name, (stack, spill, regs) in res.items():
        if "8k_directILb" in name:
            assert regs <= 80 and spill == 0 and stack == 0, (tag, name, stack, spill, regs)
    k = next(n for n in sass if "6k_leafILi4ELb1" in n)       # order 4, softened: the benchmark's kernel
    body = sass[k]
    bulk = "LEAF_BULK=0" not in " ".join(defs)
    assert (count(body, r"\bUBLKCP") > 0) == bulk              # cp.async.bulk global -> shared
    assert (count(body, r"SYNCS\.PHASECHK") > 0) == bulk       # mbarrier try_wait
    # LDGSTS: the per-lane 16-byte cp.async rows of the row fill (dozens); the bulk build keeps exactly one, the 8-byte prefetch of the
    # list entries for the tile after next
    assert (count(body, r"\bLDGSTS") > 1) == (not bulk) and count(body, r"\bLDGSTS") >= 1
    assert count(body, r"\bFFMA2\b|\bFADD2\b|\bFMUL2\b") == 0  # two-wide FP32 does not pay on B200 (profiles/r01o_summary.md)
    rows = 2 if "ROWS=2" in " ".join(defs) else 4
    assert count(body, r"MUFU\.RSQ") >= 16 * rows             # 16 targets x rows interactions in the unrolled tile loop
