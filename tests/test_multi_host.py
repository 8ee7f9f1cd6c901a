"""CPU tests (gloo, world_size 2) of the host-side logic of the N > 1 path: rank slices of the
synthetic workload tile the global particle set, the 128-byte communicator id travels from rank 0
to the others, the timing reduction is a MAX over ranks, and the reference arm runs on rank 0 only."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbody_b200 import workloads


def _worker(rank, world, port, n, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = n * rank // world, n * (rank + 1) // world
    P = workloads.plummer(hi - lo, start=lo, n_total=n)
    uid = [bytes(range(128)) if rank == 0 else None]       # stands in for nbody_cuda_comm_unique_id()
    dist.broadcast_object_list(uid, src=0)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, float(P[:, 0:3].sum())))
    np.save(os.path.join(tmp, f"slice{rank}.npy"), P)
    with open(os.path.join(tmp, f"meta{rank}.json"), "w") as f:
        json.dump({"uid_ok": uid[0] == bytes(range(128)), "tmax": float(t), "gathered": gathered}, f)
    dist.barrier()
    dist.destroy_process_group()


def test_rank_slices_tile_the_workload(tmp_path):
    n, world = 10001, 2
    mp.spawn(_worker, args=(world, 29531, n, str(tmp_path)), nprocs=world, join=True)
    full = workloads.plummer(n)
    parts = [np.load(tmp_path / f"slice{r}.npy") for r in range(world)]
    assert np.array_equal(np.concatenate(parts, axis=0), full)
    for r in range(world):
        meta = json.load(open(tmp_path / f"meta{r}.json"))
        assert meta["uid_ok"] and meta["tmax"] == float(world)
        assert [g[:2] for g in meta["gathered"]] == [[n * k // world, n * (k + 1) // world] for k in range(world)]


def test_uniform_slices_and_force_constant():
    a = workloads.uniform_cube(1000)
    b = np.concatenate([workloads.uniform_cube(400, start=0), workloads.uniform_cube(600, start=400)])
    assert np.array_equal(a, b)
    assert workloads.force_constant("plummer", 1000) == 1.0
    assert abs(workloads.force_constant("uniform", 1000) * a[:, 8].sum() - 1.0) < 0.05


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--cpu-sample", "256"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                        "--warmup", "1", "--cpu-sample", "512"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "particle-steps/s"
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def test_cpp_wrapper_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = str(tmp_path / "nbody_main")
    import nbody_b200
    if not os.path.exists(nbody_b200.LIB_PATH):
        nbody_b200.build_library()
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "nbody_main.cpp"), "-L" + os.path.join(ROOT, "nbody_b200"), "-lnbody_cuda",
                           "-Wl,-rpath," + os.path.join(ROOT, "nbody_b200"), "-o", exe])
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu tests")
    r = subprocess.run([exe, "--n", "64", "--steps", "1", "--quiet", "--csv", "none"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_rebalance_rule():
    """Per-step load rebalancing, host side (nbody_b200/csrc/balance.h through the C ABI; no device involved)."""
    import ctypes as C
    import nbody_b200
    if not os.path.exists(nbody_b200.LIB_PATH):
        nbody_b200.build_library()
    L = C.CDLL(nbody_b200.LIB_PATH)

    def rebalance(part, work, damping):
        world = len(work)
        out = (C.c_uint32 * (world + 1))()
        rc = L.nbody_cuda_rebalance(world, (C.c_uint32 * (world + 1))(*part), (C.c_float * world)(*work), C.c_float(damping), out)
        assert rc == 0
        return list(out)

    n = 1 << 20
    even = [n * r // 4 for r in range(5)]
    assert rebalance(even, [10.0, 10.0, 10.0, 10.0], 0.5) == even              # balanced: nothing moves
    assert rebalance(even, [10.0, 0.0, 10.0, 10.0], 0.5) == even               # unusable timing: keep the partition
    assert rebalance(even, [10.0, 12.0, 9.0, 11.0], 0.0) == even               # damping 0: keep the partition
    slow0 = rebalance(even, [20.0, 10.0, 10.0, 10.0], 1.0)                     # rank 0 is twice as slow: its slice shrinks
    assert slow0[0] == 0 and slow0[4] == n and slow0 == sorted(slow0)
    assert abs(slow0[1] - n * 0.25 * (12.5 / 20.0)) <= 1                        # 12.5 of its 20 units of work stay
    assert abs(slow0[2] - (n // 4 + n * 0.25 * 0.5)) <= 1                       # total 50: the second cut is 5 units into rank 1
    half = rebalance(even, [20.0, 10.0, 10.0, 10.0], 0.5)
    assert abs(half[1] - (even[1] + slow0[1]) / 2) <= 1                         # damping: half of the correction per step
    # iterating the rule on a fixed cost density converges to equal work (cost 4 per particle in the first quarter, 1 elsewhere)
    def work_of(part):
        cum = lambda x: 4.0 * min(x, n // 4) + max(0, x - n // 4)
        return [cum(part[r + 1]) - cum(part[r]) for r in range(4)]
    part = even
    for _ in range(12):
        part = rebalance(part, work_of(part), 0.5)
    w = work_of(part)
    assert max(w) / (sum(w) / 4) - 1.0 < 0.01
    # degenerate inputs
    assert rebalance([0, 7], [3.0], 0.5) == [0, 7]
    two = rebalance([0, 0, 100], [1.0, 9.0], 1.0)                              # an empty slice that reported time receives work
    assert two[0] == 0 and two[2] == 100 and 0 < two[1] <= 100


def _dist_sort_worker(rank, world, port, tmp):
    """The data flow of NBODY_FLAG_DIST_SORT with gloo standing in for NCCL: keys of the own slice, slice sort, all-gather of the
    sorted (key, index) runs with unequal counts, then the merge rounds of merge_path.h (host build) on every rank."""
    import ctypes as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    keys = np.load(os.path.join(tmp, "keys.npy"))                      # the replicated state every rank holds (previous order)
    part = np.load(os.path.join(tmp, "part.npy"))                      # the previous step's slice boundaries, known to every rank
    lo, hi = int(part[rank]), int(part[rank + 1])
    order = np.argsort(keys[lo:hi], kind="stable")                     # the rank's radix sort of its own slice
    my_k, my_i = keys[lo:hi][order], (np.arange(lo, hi, dtype=np.uint32))[order]
    gk, gi = [None] * world, [None] * world
    dist.all_gather_object(gk, my_k)                                   # = one grouped broadcast per rank (comm.cu: exchange)
    dist.all_gather_object(gi, my_i)
    k, i = np.ascontiguousarray(np.concatenate(gk)), np.ascontiguousarray(np.concatenate(gi))
    L = C.CDLL(os.path.join(ROOT, "tests", "host", "libmerge_host.so"))
    L.merge_all_runs.argtypes = [np.ctypeslib.ndpointer(np.uint64, flags="C"), np.ctypeslib.ndpointer(np.uint32, flags="C"), C.c_uint64,
                                 np.ctypeslib.ndpointer(np.uint32, flags="C"), C.c_int]
    L.merge_all_runs(k, i, len(k), np.ascontiguousarray(part, np.uint32), world)
    np.save(os.path.join(tmp, f"perm{rank}.npy"), i)
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_sort_data_flow_across_two_processes(tmp_path):
    src = os.path.join(ROOT, "tests", "host", "merge_host.cpp")
    so = os.path.join(ROOT, "tests", "host", "libmerge_host.so")
    hdr = os.path.join(ROOT, "nbody_b200", "csrc", "merge_path.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-Werror", src, "-o", so])
    rng = np.random.default_rng(5)
    n, world = 30011, 2
    keys = rng.integers(0, 1 << 20, n, dtype=np.uint64)                # plenty of ties
    part = np.array([0, 13007, n], np.int64)                           # unequal slices
    np.save(tmp_path / "keys.npy", keys)
    np.save(tmp_path / "part.npy", part)
    mp.spawn(_dist_sort_worker, args=(world, 29533, str(tmp_path)), nprocs=world, join=True)
    ref = np.argsort(keys, kind="stable")
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"perm{r}.npy").astype(np.int64), ref)   # every rank: the stable sort of all keys
