"""The experimental leaf-kernel variants must keep compiling for sm_100a within the leaf kernel's register
budget and without spills, and must really contain what they are about: TMA bulk copies with mbarrier completion
(-DNBODY_LEAF_BULK=1: UBLKCP / SYNCS in the SASS), two-wide FP32 instructions (-DNBODY_P2P_F32X2=1: FADD2 / FMUL2 / FFMA2).
The default build must contain none of them: it is the library every round-1 measurement was taken with."""
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


@pytest.mark.parametrize("tag,defs", [("default", []), ("bulk", ["-DNBODY_LEAF_BULK=1"]), ("x2", ["-DNBODY_P2P_F32X2=1"]),
                                      ("bulk_x2", ["-DNBODY_LEAF_BULK=1", "-DNBODY_P2P_F32X2=1"])])
def test_variant_builds_within_budget_and_contains_its_instructions(tmp_path, tag, defs):
    res, sass = compile_leaf(tmp_path, tag, defs)
    leaf = {k: v for k, v in res.items() if "6k_leafILi" in k}
    assert len(leaf) == 6                                     # orders 2, 3, 4 x softened / unsoftened
    for name, (stack, spill, regs) in leaf.items():
        assert regs <= 128 and spill == 0 and stack == 0, (tag, name, stack, spill, regs)
    for name, (stack, spill, regs) in res.items():
        if "8k_directILb" in name:
            assert regs <= 80 and spill == 0 and stack == 0, (tag, name, stack, spill, regs)
    k = next(n for n in sass if "6k_leafILi4ELb1" in n)       # order 4, softened: the benchmark's kernel
    body = sass[k]
    bulk, packed = "LEAF_BULK" in " ".join(defs), "F32X2" in " ".join(defs)
    assert (count(body, r"\bUBLKCP") > 0) == bulk              # cp.async.bulk global -> shared
    assert (count(body, r"SYNCS\.PHASECHK") > 0) == bulk       # mbarrier try_wait
    assert (count(body, r"\bLDGSTS") > 0) == (not bulk)        # the per-lane 16-byte cp.async rows
    n2 = count(body, r"\bFFMA2\b")
    if packed:
        # 8 target pairs x 4 rows x (3 FADD2 + 6 FFMA2 + 3 FMUL2) and 2 MUFU.RSQ per packed interaction
        assert n2 == 192 and count(body, r"\bFADD2\b") == 96 and count(body, r"\bFMUL2\b") == 96
        assert count(body, r"MUFU\.RSQ") >= 64
        d = next(n for n in sass if "8k_directILb1" in n)
        assert count(sass[d], r"\bFFMA2\b") > 0
    else:
        assert n2 == 0 and count(body, r"\bFADD2\b") == 0 and count(body, r"\bFMUL2\b") == 0


def test_pair_m2l_variant_builds_and_is_two_wide(tmp_path):
    """-DNBODY_M2L_PAIR=1 (two sibling targets per warp): fits 3 CTAs of 128 threads per SM (170 registers) with at most a few
    bytes of spill, and its arithmetic is FFMA2 / FMUL2; the default m2l.o has no two-wide instruction."""
    src = os.path.join(ROOT, "nbody_b200", "csrc", "m2l.cu")
    obj = str(tmp_path / "m2l_pair.o")
    r = subprocess.run([NVCC, "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--expt-relaxed-constexpr",
                        "-Xptxas", "-v", "-DNBODY_M2L_PAIR=1", "-c", src, "-o", obj], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = {"pair": (r.stdout + r.stderr, subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout)}
    import nbody_b200
    default_obj = os.path.join(os.path.dirname(nbody_b200.LIB_PATH), "build", "m2l.o")     # the product build (nbody_b200/build.py)
    if os.path.exists(default_obj):
        sass_default = subprocess.run(["cuobjdump", "-sass", default_obj], capture_output=True, text=True).stdout
        assert not re.search(r"\bFFMA2\b|\bFMUL2\b|\bFADD2\b", sass_default)
    log, sass = out["pair"]
    found = 0
    for m in re.finditer(r"Compiling entry function '(\S*k_m2l_pair\S*)' for 'sm_100a'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads"
                         r".*?Used (\d+) registers", log, re.S):
        found += 1
        assert int(m.group(5)) <= 170 and int(m.group(3)) + int(m.group(4)) <= 64, m.groups()
    assert found == 3
    body, name = [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
        elif name and "k_m2l_pairILi4E" in name:
            body.append(line)
    assert count(body, r"\bFFMA2\b") >= 250 and count(body, r"\bFMUL2\b") >= 80
