"""CPU guard on what ptxas allocated for the hot kernels (nbody_b200/build/*.o.log, written by nbody_b200/build.py with
-Xptxas -v): the occupancy each kernel was tuned for (profiles/r01k_summary.md) depends on these register
counts, and a silent spill in the interaction loops would cost far more than any test would notice functionally."""
import os
import re

import pytest

import nbody_b200

BUILD = os.path.join(os.path.dirname(nbody_b200.LIB_PATH), "build")


def kernels():
    if not os.path.isdir(BUILD) or not any(f.endswith(".o.log") for f in os.listdir(BUILD)):
        nbody_b200.build_library(force=True)
    out = {}
    for f in sorted(os.listdir(BUILD)):
        if not f.endswith(".o.log"):
            continue
        txt = open(os.path.join(BUILD, f)).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads"
                             r".*?Used (\d+) registers", txt, re.S):
            out[m.group(1)] = {"stack": int(m.group(2)), "spill": int(m.group(3)) + int(m.group(4)), "regs": int(m.group(5)), "file": f}
    return out


K = kernels()


def select(pattern):
    sel = {k: v for k, v in K.items() if re.search(pattern, k) and "cub" not in k}
    assert sel, pattern
    return sel


def test_every_kernel_is_built_for_sm_100a_and_found():
    names = "".join(K)
    for kernel in ("k_keys", "k_sort_hist", "k_sort_scatter", "k_gather", "k_level_count", "k_level_split", "k_p2m", "k_m2m", "k_traverse",
                   "k_m2l", "k_l2l", "k_leaf", "k_direct", "k_acc_max", "k_partition", "k_gather_vel", "k_keys_range", "k_merge_partition", "k_merge_runs"):
        assert kernel in names, kernel


@pytest.mark.parametrize("pattern,max_regs,why", [
    (r"6k_leafILi\d", 128, "4 CTAs of 128 threads per SM (launch bounds), the configuration of every leaf-kernel measurement"),
    (r"5k_m2lILi[234]ELi8E", 128, "2 CTAs of 256 threads per SM"),
    (r"5k_m2lILi5ELi8E", 255, "order 5 (56 coefficients): one CTA of 256 threads per SM, no spills"),
    (r"10k_traverse", 48, "5 CTAs of 256 threads per SM: the traversal is occupancy-sensitive (15 % from 4 -> 5 CTAs)"),
    (r"8k_directILb", 80, "3 CTAs of 256 threads per SM"),
])
def test_hot_kernels_keep_their_register_budget_without_spills(pattern, max_regs, why):
    for name, r in select(pattern).items():
        assert r["regs"] <= max_regs, (name, r, why)
        assert r["spill"] == 0 and r["stack"] == 0, (name, r, "spill in a hot loop")


def test_no_other_kernel_of_ours_spills_beyond_the_known_one():
    for name, r in K.items():
        if "cub" in name:
            continue
        if "k_sort_scatter" in name:     # 16 keys + packed ranks per thread at 3 CTAs/SM: 40 bytes of spill, accepted and measured
            assert r["regs"] <= 80 and r["spill"] <= 96, (name, r)
        else:
            assert r["spill"] == 0, (name, r)
