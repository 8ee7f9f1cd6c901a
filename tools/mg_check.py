"""Multi-GPU check (run under torchrun, one rank per GPU): the distributed step must reproduce the
single-GPU step — same tree-ordered state on every rank — and the owned ranges must tile [0, N)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import nbody_b200
from nbody_b200 import workloads

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = sys.argv[3] if len(sys.argv) > 3 else "plummer"
eta = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0  # > 0: variable time step (the maximum |a| is all-gathered across ranks)
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0     # nbody_cuda_config.flags of the DISTRIBUTED run (e.g. 64 = NBODY_FLAG_DIST_SORT); the 1-GPU reference runs with 0
lo, hi = n * rank // world, n * (rank + 1) // world
P = workloads.plummer(hi - lo, start=lo, n_total=n) if kind == "plummer" else workloads.GENERATORS[kind](n)[lo:hi]
uid = [nbody_b200.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
sim = nbody_b200.CudaSimulation([1, 1, 1], P, 1e-3, device=local, time_step_eta=eta, flags=flags,
                                _distributed={"unique_id": uid[0], "n_global": n, "global_offset": lo, "rank": rank, "world": world})
ranges = []
dts = []
t0 = time.time()
for k in range(steps):
    sim.step()
    dts.append(sim.time_step())
    st_k = sim.stats()
    if rank == 0:
        print(f"step {k}: dt {dts[-1]['last']:.3e} owned {sim.owned_range()} work imbalance {st_k['work_imbalance']:.3f} ms_total {st_k['ms_total']:.2f}", flush=True)
torch.cuda.synchronize()
dt = time.time() - t0
out = sim.particles()
first, count = sim.owned_range()
allr = [None] * world
dist.all_gather_object(allr, (first, count, sim.stats()["ms_total"], sim.stats()["p2p_interactions"], sim.stats()["m2l_interactions"], sim.time_step()["next"]))
ok = True
if rank == 0:
    print("owned ranges:", [(a, b) for a, b, *_ in allr], "ms_total per rank:", [round(x[2], 2) for x in allr], flush=True)
    ok &= all(x[5] == allr[0][5] for x in allr)  # every rank derived the same next time step, bit for bit
    cover = sorted((a, a + b) for a, b, *_ in allr)
    ok &= cover[0][0] == 0 and cover[-1][1] == n and all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
    ref = nbody_b200.CudaSimulation([1, 1, 1], workloads.GENERATORS[kind](n), 1e-3, device=local, time_step_eta=eta)
    ref_dts = []
    for _ in range(steps):
        ref.step()
        ref_dts.append(ref.time_step())
    if eta > 0:
        agree = all(abs(a["last"] / b["last"] - 1) < 1e-4 and abs(a["next"] / b["next"] - 1) < 1e-4 and abs(a["acc_max"] / b["acc_max"] - 1) < 1e-4
                    for a, b in zip(dts, ref_dts)) and dts[-1]["next"] < 1e-3
        print(f"time steps agree: {agree}  multi {[round(d['last'], 8) for d in dts]} single {[round(d['last'], 8) for d in ref_dts]} max|a| {dts[-1]['acc_max']:.4g}", flush=True)
        ok &= agree
    r = ref.particles()
    same_perm = np.array_equal(ref.permutation(), sim.permutation())
    dpos = np.abs(r[:, 0:3] - out[:, 0:3]).max(); dvel = np.abs(r[:, 4:7] - out[:, 4:7]).max()
    st = ref.stats()
    print(f"single-GPU vs {world}-GPU after {steps} steps: same order {same_perm}, max|dx| {dpos:.3e}, max|dv| {dvel:.3e}; "
          f"P2P evals single {st['p2p_interactions']} vs sum {sum(x[3] for x in allr)}; M2L single {st['m2l_interactions']} vs sum {sum(x[4] for x in allr)}; "
          f"single ms {st['ms_total']:.2f}", flush=True)
    if kind == "plummer":  # an equilibrium model: 2-GPU and 1-GPU runs stay together; a cold two-galaxy collision is chaotic
        ok &= same_perm and dpos < 1e-5 and dvel < 1e-3 and st["p2p_interactions"] == sum(x[3] for x in allr)
# every rank holds the same full state
h = torch.tensor([float(np.abs(out[:, 0:7]).sum())], device="cuda", dtype=torch.float64)
hs = [torch.zeros_like(h) for _ in range(world)]
dist.all_gather(hs, h)
if rank == 0:
    same = all(abs(float(x) - float(hs[0])) < 1e-6 * abs(float(hs[0])) for x in hs)
    print("state identical on all ranks:", same, flush=True)
    ok &= same
    print("MG_CHECK", "PASS" if ok else "FAIL", flush=True)
sim.close()
dist.destroy_process_group()
