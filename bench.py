#!/usr/bin/env python
"""bench.py — particle-steps/s of one full FMM gravity step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torchrun, one rank per GPU)

A "step" is one Simulation::step() on the metric's configuration: a Plummer sphere
of N = 16M particles (BASELINE.json configs[2]), at the reference's physics constants
(MAC 0.5, softening 0.01). The octree node capacity — hard-coded to 8 in the reference
with a "should be adjustable" FIXME (src/open_cl_simulation.cpp:41-47) — is a tuning
parameter here: the headline runs at 48, and the same workload at the reference's 8 is
measured in the same run and reported under "reference_capacity". One JSON line on
stdout (rank 0).
  value     whole-job particle-steps/s, state resident in HBM: N * K / wall time of K blocking step() calls between
            barrier + synchronize pairs, max over ranks; device_ms_per_step is the same interval seen by CUDA events
            on the solver's own stream (max over ranks) and must agree with ms_per_step
  e2e       the same through the C ABI with HOST buffers: set_particles (H2D) +
            step + get_particles (D2H) inside the timed region
  roofline  dominant kernel vs the FP32 FMA peak (this path is FP32-pipe bound by
            design, not HBM/tensor: BASELINE.md section 3), live CUDA-event times
  cpu_baseline  the UNMODIFIED reference naive_simulation.cpp (oracle/_ref) timed on
            this box's host cores on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec at N=16M Plummer (1/2/4/8 B200); P2P FP32 TFLOP/s"
UNIT = "particle-steps/s"
N_SM = 148


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback (B200_PROFILING.md)"}


def measured_traffic(kernel, args):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture of exactly this workload (profiles/r02_traffic.json); None for any other workload."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        w = d["workload"]
        if (w["kind"], w["n"], w["order"], w["leaf_capacity"]) == (args.workload, args.n, args.order, args.leaf_capacity) and args.gpus == 1:
            return d["dram_bytes_per_launch"].get(kernel)
    except Exception:
        pass
    return None


def fp32_peak_tflops(sm_mhz):
    return N_SM * 128 * 2 * sm_mhz * 1e6 / 1e12


def m2l_flops_per_interaction(P):
    """Exact operation count of one directed M2L as implemented (expansion.cuh):
    derivative tensor (monomials, radial scalars, Hermite-form sums) + the field-only contraction."""
    mi = [(i, j, o - i - j) for o in range(P + 1) for i in range(o, -1, -1) for j in range(o - i, -1, -1)]
    nc = len(mi)
    flops = (nc - 1)            # monomials: one multiply each
    flops += 3 + 2 * 3 + 2      # displacement (3 sub), R2 (3 FMA), rsqrt (counted 2)
    flops += 1 + 2 * P          # inv2, g_k recurrences (2 mul each)
    for (i, j, k) in mi:        # D_n = sum_j c g x^(n-2j): one FMA per term (+ a constant*g product shared; ignored)
        terms = (i // 2 + 1) * (j // 2 + 1) * (k // 2 + 1)
        flops += 2 * terms
    contraction = sum(1 for n in mi if sum(n) >= 1 for m in mi if sum(n) + sum(m) <= P)
    return flops + 2 * contraction, contraction


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [x.strip() for x in line.split(",")])

    def mark(self):
        """Start of the timed region: samples before this were taken under the (identical) warm-up load."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        timed = [r for r in self.rows if r[0] >= t_mark]
        # a timed region of a few steps at 8 GPUs lasts ~0.1 s: when it holds fewer than 3 samples, the samples of the warm-up steps
        # (the same step, the same load, immediately before) are used as well, and the line says so
        use, window = (timed, "timed region") if len(timed) >= 3 else (self.rows, "warm-up + timed region (timed region too short for 3 samples)")
        sm, mx, reasons = [], [], set()
        for r in use:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_in_timed_region": len(timed), "window": window, "reasons": sorted(reasons)}


FLAG_PARTITIONED = 128  # nbody_b200.FLAG_PARTITIONED / NBODY_FLAG_PARTITIONED (include/nbody_cuda.h)


def arm_config(args, world):
    """The benchmark configuration, printed under "config" by BOTH arms: the bench contract has the reference arm run "on your
    arm's config" (a bounded sample of that workload per step; what the sample was is said in cpu_baseline.sample and
    reference_sample). Everything here follows from the command line and the rank count alone."""
    partitioned = world > 1 and args.scheme != "replicated"
    return {"workload": f"{args.workload} sphere N={args.n}" if args.workload == "plummer" else f"{args.workload} N={args.n}",
            "order": args.order, "leaf_capacity": args.leaf_capacity, "mac_ratio": 0.5, "softening": 0.01,
            "integrator": "kick-drift", "l2_policy": "working set (>1 GB of lists and particle state per step) exceeds the 126 MB L2",
            "partition": ("morton-range, partitioned state + locally essential tree" if partitioned else
                          "morton-range, replicated state" if world > 1 else "single"),
            "flags": args.flags | (FLAG_PARTITIONED if partitioned else 0)}


def reference_arm(args, rank, emit):
    """--impl reference: the unmodified reference CPU path (naive_simulation.cpp via oracle/_ref;
    the oracle port if the reference was never compiled) on a bounded sample of the same workload."""
    if rank != 0:
        return
    import oracle
    from nbody_b200 import workloads
    ns = args.cpu_sample
    P = workloads.GENERATORS[args.workload](ns, n_total=args.n) if args.workload == "plummer" else workloads.GENERATORS[args.workload](ns)
    have_ref = oracle.ref_lib() is not None
    run = (lambda steps: oracle.ref_naive_run(P, 1.0, args.dt, steps)) if have_ref else \
          (lambda steps: oracle.naive_step_as_written(P, 1.0, args.dt, steps))
    for _ in range(min(args.warmup, 1)):
        run(1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(1)
    dt = time.perf_counter() - t0
    value = ns * args.steps / dt
    pairs_per_s = ns * (ns - 1) / 2 * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": arm_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
        "reference_sample": {"n": ns, "what": f"each step = NaiveSimulation::step() on the first {ns} particles of the workload named in config "
                                              "(the solver fields of config describe the measured arm; the reference's direct sum has none of them)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference" if have_ref else "port",
                         "sample": f"NaiveSimulation::step() x{args.steps} on the first {ns} particles of the workload "
                                   f"(O(N^2), single-threaded by construction; {pairs_per_s:.3e} pair-interactions/s; "
                                   f"at the full N={args.n} the same code would deliver {2 * pairs_per_s / max(args.n - 1, 1):.3e} particle-steps/s)",
                         "host_cores_visible": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    # BASELINE config 1 exactly: the unmodified NaiveSimulation on ALL 4096 particles of the uniform cube, 10 steps
    P1 = workloads.uniform_cube(4096)
    G1 = workloads.force_constant("uniform", 4096)
    run1 = (lambda steps: oracle.ref_naive_run(P1, G1, args.dt, steps)) if have_ref else (lambda steps: oracle.naive_step_as_written(P1, G1, args.dt, steps))
    t0 = time.perf_counter()
    run1(10)
    dt1 = time.perf_counter() - t0
    line["config1"] = {"workload": "uniform cube N=4096, 10 steps (BASELINE configs[0])", "value": 4096 * 10 / dt1, "unit": UNIT,
                       "ms_per_step": 1e2 * dt1, "cores": 1, "kind": "reference" if have_ref else "port", "same_config_as_ours_config1": True}
    # For the record, the reference's FMM path itself on the same configuration: its OpenCL kernels compiled for the host
    # (oracle/_ref/libclref.so), one thread, on the oracle's octree (glade is not in the reference tree). What an OpenCL CPU device
    # would execute for OpenClSimulation::step(); its direct sum above is the faster of the two at these sizes, hence the headline.
    try:
        if oracle.clref_lib() is not None:
            keys = oracle.morton_keys(P1[:, 0:3], (1.0, 1.0, 1.0))
            t0 = time.perf_counter()
            sk, perm = oracle.sort_keys(keys)
            tree = oracle.Tree(sk, (1.0, 1.0, 1.0), 8, 21)
            _, pairs = oracle.ClRef(tree, P1[perm]).step(args.dt, repair=True, node_local_size=32)
            dtf = time.perf_counter() - t0
            line["config1"]["reference_fmm_kernels_on_host"] = {
                "value": 4096 / dtf, "unit": UNIT, "ms_per_step": 1e3 * dtf, "steps": 1, "cores": 1, "interaction_pairs": pairs,
                "what": "src/{moment,interaction,field,force}.cl compiled for the host, host defects D5/D7 repaired, octree by the oracle; "
                        "2-7 % RMS from direct summation (profiles/r02z_reference_kernels_on_host.md)"}
    except Exception as exc:  # noqa: BLE001 — informational only
        line["config1"]["reference_fmm_kernels_on_host"] = {"unavailable": str(exc)[:120]}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="plummer", choices=["plummer", "uniform", "two_galaxies"])
    ap.add_argument("--n", "--particles", dest="n", type=int, default=1 << 24)
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--leaf-capacity", type=int, default=48,
                    help="octree node capacity; the reference hard-codes 8 with a FIXME (src/open_cl_simulation.cpp:41-47), 48 is the B200 tuning")
    ap.add_argument("--no-reference-capacity", action="store_true", help="skip the additional measurement at the reference's capacity 8")
    ap.add_argument("--dt", type=float, default=1e-3)
    ap.add_argument("--cpu-sample", type=int, default=16384)
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pool-scale", type=float, default=1.0)
    ap.add_argument("--flags", type=int, default=0, help="nbody_cuda_config.flags (e.g. 32 = NBODY_FLAG_NO_OVERLAP, for A/B runs)")
    ap.add_argument("--scheme", default="auto", choices=["auto", "partitioned", "replicated"],
                    help="multi-GPU scheme: partitioned = own particles + locally essential tree (NBODY_FLAG_PARTITIONED, the default for "
                         "N > 1), replicated = round 1's replicated state and tree")
    ap.add_argument("--tau", type=float, default=None, help="low_order_tau (adaptive-order M2L); default: the library's 0.13")
    ap.add_argument("--no-accuracy", action="store_true", help="skip the accuracy check of the benched configuration")
    ap.add_argument("--accuracy-targets", type=int, default=65536)
    ap.add_argument("--slack-pct", type=int, default=0, help="partitioned mode: nbody_cuda_config.partition_slack_pct (0 = automatic)")
    ap.add_argument("--no-multi-check", action="store_true", help="N > 1: skip the comparison with a single-GPU run at N = 2^20")
    ap.add_argument("--no-config1", action="store_true", help="skip BASELINE config 1 (uniform N = 4096, 10 steps) measured alongside")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line, the JSON: anything a library prints there (NCCL's version banner under torchrun)
    # goes to stderr instead; the line itself is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        reference_arm(args, rank, emit)
        return

    import torch
    import torch.distributed as dist
    import nbody_b200
    from nbody_b200 import workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_cpus(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.n
    # this rank's slice of the global particle set (weak in memory, strong in work: N is fixed)
    lo = n * rank // world
    hi = n * (rank + 1) // world
    big = hi - lo > (1 << 26)                    # config 5 (2^27 particles per rank): no pinned staging, no end-to-end leg
    host = torch.empty((hi - lo, 12), dtype=torch.float32)
    if not big:
        host = host.pin_memory()
    Pn = host.numpy()
    for c0 in range(lo, hi, 1 << 23):            # in chunks: the generator's temporaries stay small at 2^27 particles per rank
        Pn[c0 - lo:min(c0 + (1 << 23), hi) - lo] = workloads.generate(args.workload, n, c0, min(1 << 23, hi - c0))
    partitioned = world > 1 and args.scheme != "replicated"
    base_flags = args.flags | (nbody_b200.FLAG_PARTITIONED if partitioned else 0)

    def make_sim(capacity, particles=None, n_total=None, offset=None, flags=None, kind=None):
        particles = Pn if particles is None else particles
        n_total = n if n_total is None else n_total
        cfgkw = dict(order=args.order, leaf_capacity=capacity, device=local_rank, pool_scale=args.pool_scale,
                     flags=base_flags if flags is None else flags, force_constant=workloads.force_constant(kind or args.workload, n_total))
        if args.tau is not None:
            cfgkw["low_order_tau"] = args.tau
        if args.slack_pct:
            cfgkw["partition_slack_pct"] = args.slack_pct
        if world == 1:
            return nbody_b200.CudaSimulation([1.0, 1.0, 1.0], particles, args.dt, **cfgkw)
        uid = [nbody_b200.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        return nbody_b200.CudaSimulation([1.0, 1.0, 1.0], particles, args.dt, _distributed={
            "unique_id": uid[0], "n_global": n_total, "global_offset": lo if offset is None else offset, "rank": rank, "world": world}, **cfgkw)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(sim, warm, k, sampler=None):
        """`warm` untimed steps, then exactly k steps between barriers; returns (seconds [max over ranks], per-step stage ms, counts)."""
        if sampler is not None:
            sampler.start()
            time.sleep(0.4)  # nvidia-smi needs a moment to produce its first sample
        for _ in range(warm):
            sim.step()
        barrier()
        if sampler is not None:
            sampler.mark()
        sums, cnts, imb = {}, {}, []
        t0 = time.perf_counter()
        for _ in range(k):
            sim.step()
            for key, v in sim.stats().items():
                if key.startswith("ms_"):
                    sums[key] = sums.get(key, 0.0) + v
                else:
                    cnts[key] = v
            imb.append(round(cnts.get("work_imbalance", 0.0), 4))
        barrier()
        dt = time.perf_counter() - t0
        dev_ms = sums.get("ms_total", 0.0)  # CUDA events on the solver's stream, first launch to last of every step
        if world > 1:
            tt = torch.tensor([dt, dev_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt, dev_ms = float(tt[0].item()), float(tt[1].item())
        cnts["device_ms_per_step"] = dev_ms / k
        cnts["work_imbalance_per_step"] = imb
        return dt, {key: v / k for key, v in sums.items()}, cnts

    sim = make_sim(args.leaf_capacity)
    sampler = ClockSampler(local_rank)
    elapsed, stage_ms, counts = timed_steps(sim, max(args.warmup, 3), args.steps, sampler)
    clocks = sampler.stop()
    if world > 1:  # every rank samples its own GPU: eight GPUs under load at once need not hold the clock one GPU holds alone
        allc = [None] * world
        dist.all_gather_object(allc, {"sm_mhz": clocks["sm_mhz"], "samples": clocks["samples"], "reasons": clocks["reasons"]})
        clocks["per_rank"] = allc
        known = [c["sm_mhz"] for c in allc if c["sm_mhz"] is not None]
        if known:
            clocks["sm_mhz_min_over_ranks"] = min(known)
        clocks["reasons"] = sorted(set(r for c in allc for r in c["reasons"]))
    value = n * args.steps / elapsed
    K = args.steps
    rank_ms = None
    if world > 1:
        RANK_COLS = ("ms_sort", "ms_tree", "ms_upsweep", "ms_traverse", "ms_m2l", "ms_l2l", "ms_leaf", "ms_comm", "ms_import", "ms_halo", "ms_balance", "ms_total")
        mine = torch.tensor([stage_ms.get(k, 0.0) for k in RANK_COLS], device="cuda", dtype=torch.float64)
        allm = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allm, mine)
        rank_ms = [[round(float(x), 2) for x in t] for t in allm]

    # ---- end to end through the C ABI with host buffers --------------------------------
    # every step: H2D of this rank's slice of the state from pinned host memory (set_owned_particles, which
    # also re-assembles the full state over NVLink when N > 1), the step, D2H of the rank's updated slice.
    first, count = sim.owned_range()
    e2e_steps = 0 if big else args.e2e_steps
    # owned counts drift with the per-step rebalancing (a few leaves on the Plummer benchmark): room for +100 % on a multi-GPU run
    out = torch.empty(((max(count, 1) + ((n // world) if world > 1 else 0) + 1024) if e2e_steps else 16, 12), dtype=torch.float32).pin_memory()
    if e2e_steps:
        sim.owned_particles_into_ptr(out.data_ptr(), out.shape[0])  # warm the export path
    barrier()
    h2d = d2h = 0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        first, count = sim.owned_range()
        sim.set_owned_particles_ptr(out.data_ptr(), count)
        h2d += 48 * count
        sim.step()
        first, count = sim.owned_range()
        sim.owned_particles_into_ptr(out.data_ptr(), out.shape[0])
        d2h += 48 * count
    barrier()
    e2e_t = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_t, float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        e2e_t, h2d, d2h = float(mx[0]), float(sm[1]), float(sm[2])
    e2e = {"value": n * e2e_steps / e2e_t, "unit": UNIT, "h2d_bytes_per_step": int(h2d / max(e2e_steps, 1)),
           "d2h_bytes_per_step": int(d2h / max(e2e_steps, 1)), "ms_per_step": 1e3 * e2e_t / max(e2e_steps, 1), "steps": e2e_steps}
    if not e2e_steps:
        e2e["skipped"] = "more than 2^26 particles per rank: the end-to-end leg would pin > 12 GB of host memory per rank"

    sim.close()

    # ---- accuracy of the benched configuration: one force evaluation of the same workload, capacity, order and
    # tau on the initial condition; rank 0 compares the accelerations of 65,536 of ITS particles with direct summation over ALL N
    # sources (the all-pairs GPU kernel, FP32 with compensated tile sums; tests/test_gpu_parity.py checks that kernel against FP64)
    accuracy = None
    if not args.no_accuracy:
        fsim = make_sim(args.leaf_capacity, flags=base_flags | nbody_b200.FLAG_NO_INTEGRATE)
        fsim.step()
        if partitioned or world == 1:
            own, acc = fsim.particles(), fsim.accelerations()
        else:
            acc = fsim.accelerations()           # collective in the replicated scheme
            f0, c0 = fsim.owned_range()
            own, acc = fsim.particles()[f0:f0 + c0], acc[f0:f0 + c0]
        fsim.close()
        if rank == 0:
            ntg = min(args.accuracy_targets, own.shape[0])
            tg = np.linspace(0, own.shape[0] - 1, ntg).astype(np.int64)
            tpos = np.ascontiguousarray(np.concatenate([own[tg, 0:3], own[tg, 9:10]], axis=1))
            f = np.zeros((ntg, 3), np.float64)
            chunk = 1 << 25                      # the sources in chunks (the field is linear in them): the full set need not fit anywhere
            for c0 in range(0, n, chunk):
                part = Pn[c0:c0 + chunk] if world == 1 else workloads.generate(args.workload, n, c0, min(chunk, n - c0))
                src = np.ascontiguousarray(np.concatenate([part[:, 0:3], part[:, 9:10]], axis=1))
                f += nbody_b200.direct_field(src, tpos, 0.01, device=local_rank)[0]
            ref = f * (workloads.force_constant(args.workload, n) * own[tg, 9] / own[tg, 8])[:, None]
            err = float(np.sqrt(((acc[tg].astype(np.float64) - ref) ** 2).sum() / (ref ** 2).sum()))
            accuracy = {"rms_rel": err, "targets": int(ntg), "sources": int(n), "bar": 1e-3,
                        "what": "accelerations of one force evaluation of this configuration vs direct summation over all sources (GPU all-pairs kernel)"}
        barrier()

    # ---- N > 1: the multi-GPU step against a single-GPU run (same workload at N = 2^20, 3 steps): same tree order, same
    # trajectories to FP32 round-off, the P2P work of the ranks sums to the single-GPU count exactly
    multi_check = None
    if world > 1 and not args.no_multi_check:
        multi_check = multi_gpu_check(args, rank, world, local_rank, make_sim, partitioned, dist, torch)

    # ---- BASELINE config 1 (the reference's own CPU-runnable case): uniform cube N = 4096, 10 steps, measured alongside so that one
    # same-configuration ratio against --impl reference exists
    config1 = None
    if world == 1 and rank == 0 and not args.no_config1:
        config1 = config1_ours(args, local_rank)
    ref_cap = None
    if args.leaf_capacity != 8 and not args.no_reference_capacity:
        # the same workload at the reference's hard-coded node capacity (src/open_cl_simulation.cpp:41-47), for the record
        sim8 = make_sim(8)
        dt8, st8, c8 = timed_steps(sim8, 3, min(args.steps, 3))
        sim8.close()
        ref_cap = {"leaf_capacity": 8, "value": n * min(args.steps, 3) / dt8, "unit": UNIT, "ms_per_step": 1e3 * dt8 / min(args.steps, 3), "device_ms_per_step": c8.pop("device_ms_per_step", None),
                   "stage_ms": st8, "m2l_interactions": c8.get("m2l_interactions"), "p2p_interactions": c8.get("p2p_interactions")}
    p2p_micro = None
    if rank == 0:
        # P2P FP32 microbenchmark (BASELINE metric, second half): the all-pairs tiled kernel on 2^18 x 2^20 bodies of the workload
        nm = min(Pn.shape[0], 1 << 20)
        ntg = min(nm, N_SM * 2 * 1024)  # 2 CTAs of 1024 targets per SM: whole waves
        src = np.ascontiguousarray(np.concatenate([Pn[:nm, 0:3], Pn[:nm, 9:10]], axis=1))
        _, ms = nbody_b200.direct_field(src, src[:ntg], 0.01, device=local_rank, repeats=3)
        p2p_micro = {"kernel": "k_direct", "targets": ntg, "sources": nm, "ms": ms, "tflops": 20.0 * ntg * nm / (ms * 1e-3) / 1e12}

    if rank == 0:
        peaks = measured_peaks()
        peak = fp32_peak_tflops(peaks["sm_max_mhz"])
        m2l_flops, m2l_fma = m2l_flops_per_interaction(args.order)
        m2l_flops_lo, _ = m2l_flops_per_interaction(max(args.order - 1, 1))
        n_lo = counts.get("m2l_interactions_low", 0)
        m2l_total_flops = (counts["m2l_interactions"] - n_lo) * m2l_flops + n_lo * m2l_flops_lo
        m2l_tf = m2l_total_flops / (stage_ms["ms_m2l"] * 1e-3) / 1e12 if stage_ms.get("ms_m2l", 0) > 0 else 0.0
        p2p_tf = counts["p2p_interactions"] * 20 / (stage_ms["ms_leaf"] * 1e-3) / 1e12 if stage_ms.get("ms_leaf", 0) > 0 else 0.0
        dominant = "m2l" if stage_ms.get("ms_m2l", 0) >= stage_ms.get("ms_leaf", 0) else "p2p"
        roof = {
            "bound": "fp32", "kernel": "k_m2l (M2L, both launches)" if dominant == "m2l" else "k_leaf (P2P+L2P+integrate)",
            "achieved": m2l_tf if dominant == "m2l" else p2p_tf, "peak": peak, "unit": "TFLOP/s",
            "frac": (m2l_tf if dominant == "m2l" else p2p_tf) / peak,
            "traffic": measured_traffic("k_m2l" if dominant == "m2l" else "k_leaf", args),
            "traffic_source": "profiles/r02_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this "
                              "kernel on this workload (not measured in this run; null for any other workload or N > 1)",
            "peak_source": f"148 SM x 128 lanes x 2 x {peaks['sm_max_mhz']} MHz ({peaks['source']}; FP32 FMA peak, derived)",
            "algorithmic_flops_per_unit": (f"{m2l_flops} per order-{args.order} M2L, {m2l_flops_lo} per order-{args.order - 1} M2L "
                                           f"({n_lo} of {counts['m2l_interactions']} run at the lower order)") if dominant == "m2l" else 20,
            "units_per_step": counts["m2l_interactions"] if dominant == "m2l" else counts["p2p_interactions"],
            "kernel_ms": stage_ms["ms_m2l"] if dominant == "m2l" else stage_ms["ms_leaf"],
            "share_of_step": (stage_ms["ms_m2l"] if dominant == "m2l" else stage_ms["ms_leaf"]) / stage_ms["ms_total"],
        }
        # the other stages against their own rooflines (live stage times of rank 0; never allowed to break the line)
        try:
            n_own = int(counts.get("n_particles", n))
            sort_gbs = 72.0 * n_own / (stage_ms["ms_sort"] * 1e-3) / 1e9 if stage_ms.get("ms_sort", 0) > 0 else 0.0
            other_roofs = {
                "m2l": {"bound": "fp32", "kernel": "k_m2l (both launches)", "achieved": m2l_tf, "peak": peak, "unit": "TFLOP/s", "frac": m2l_tf / peak,
                        "kernel_ms": stage_ms.get("ms_m2l")},
                "keys_sort_gather": {"bound": "hbm", "kernel": "k_keys + 8 radix passes + k_gather", "achieved": sort_gbs, "peak": peaks["hbm_gbs"],
                                     "unit": "GB/s", "frac": sort_gbs / peaks["hbm_gbs"], "kernel_ms": stage_ms.get("ms_sort"),
                                     "algorithmic_bytes_per_unit": "72 B per particle compulsory (read 32 B state, write 32 B state, 8 B key); "
                                                                   "as implemented: 8 passes x 32 B + keys + gather"},
            }
        except Exception as exc:  # noqa: BLE001
            other_roofs = {"error": str(exc)[:120]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * elapsed / K, "device_ms_per_step": counts.pop("device_ms_per_step", None), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": arm_config(args, world), "library": os.path.basename(nbody_b200.LIB_PATH), "cpu_binding": numa,
            "clocks": clocks, "e2e": e2e, "gpu_launches": K * launches_per_step(sim, world, counts.get("n_levels"), partitioned),
            "roofline": roof, "other_rooflines": other_roofs,
            "p2p_fp32_tflops": {"tree_p2p_kernel": p2p_tf, "tree_p2p_frac_of_peak": p2p_tf / peak,
                                "all_pairs_kernel": p2p_micro["tflops"], "all_pairs_frac_of_peak": p2p_micro["tflops"] / peak,
                                "all_pairs_config": p2p_micro},
            "m2l_fp32_tflops": m2l_tf, "reference_capacity": ref_cap,
            "stage_ms": stage_ms, "counts": counts,
            "accuracy": accuracy, "multi_gpu_check": multi_check, "config1": config1,
        }
        if rank_ms is not None:
            line["per_rank_ms"] = {"columns": [c[3:] for c in RANK_COLS], "rows": rank_ms}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        emit(line)
    sim.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        if accuracy is not None and not (accuracy["rms_rel"] <= accuracy["bar"]):
            raise SystemExit(f"accuracy check failed: RMS relative error {accuracy['rms_rel']:.3e} > 1e-3")
        if multi_check is not None and not multi_check["pass"]:
            raise SystemExit(f"multi-GPU check failed: {multi_check}")


def bind_to_gpu_cpus(index):
    """One process per GPU: run on the CPU cores NVML reports as local to this GPU, so that the pinned staging buffers of the
    end-to-end leg (first touch) live on the GPU's own NUMA node — what any MPI launcher's rank binding does; torchrun does not."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if word >> b & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": f"{cpus[0]}-{cpus[-1]}", "count": len(cpus)}
    except Exception as e:  # no NVML, no permission: run unbound
        return {"unbound": str(e)[:80]}
    return None


def multi_gpu_check(args, rank, world, local_rank, make_sim, partitioned, dist, torch):
    import nbody_b200
    from nbody_b200 import workloads
    n2, steps = 1 << 20, 3
    lo, hi = n2 * rank // world, n2 * (rank + 1) // world
    sim = make_sim(args.leaf_capacity, particles=workloads.generate(args.workload, n2, lo, hi - lo), n_total=n2, offset=lo)
    sim.step()
    first = (sim.permutation() if partitioned else sim.permutation()[slice(*(lambda f, c: (f, f + c))(*sim.owned_range()))],
             sim.stats()["p2p_interactions"])
    for _ in range(steps - 1):
        sim.step()
    st = sim.stats()
    if partitioned:
        mine, perm = sim.particles(), sim.permutation()
    else:
        f0, c0 = sim.owned_range()
        full, fperm = sim.particles(), sim.permutation()
        mine, perm = full[f0:f0 + c0], fperm[f0:f0 + c0]
        h = torch.tensor([float(np.abs(full[:, 0:7]).sum(dtype=np.float64))], device="cuda", dtype=torch.float64)
        hs = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
    sim.close()
    parts = [None] * world if rank == 0 else None
    dist.gather_object((mine, perm, st["p2p_interactions"], st["m2l_interactions"], st["device_bytes"], first[0], first[1]), parts, dst=0)
    res = None
    if rank == 0:
        got = np.concatenate([p[0] for p in parts])
        gperm = np.concatenate([p[1] for p in parts])
        ref = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], workloads.generate(args.workload, n2), args.dt, order=args.order,
                                        leaf_capacity=args.leaf_capacity, device=local_rank,
                                        force_constant=workloads.force_constant(args.workload, n2),
                                        **({"low_order_tau": args.tau} if args.tau is not None else {}))
        ref.step()
        rperm1, rp2p1 = ref.permutation(), int(ref.stats()["p2p_interactions"])
        for _ in range(steps - 1):
            ref.step()
        r, rperm, rst = ref.particles(), ref.permutation(), ref.stats()
        ref.close()
        # first step (identical input): the same tree order, the same P2P work particle for particle
        gperm1 = np.concatenate([p[5] for p in parts])
        same_order = gperm1.shape == rperm1.shape and bool(np.array_equal(gperm1, rperm1))
        p2p_sum = int(sum(p[6] for p in parts))
        # (bars: |dx| < 1e-5 of a unit box, |dv| < 2e-3 where speeds reach 8 in the core of the Plummer model; measured 8e-7 and 4.7e-4 at 8 ranks)
        # after `steps` steps: the same trajectories by particle identity (the runs differ by FP32 round-off, so a particle that sits
        # 1e-7 from a cell boundary may sort differently: the order itself is only compared on the first step)
        complete = got.shape == r.shape and bool(np.array_equal(np.sort(gperm), np.arange(n2, dtype=np.uint32)))
        ia, ib = np.argsort(gperm), np.argsort(rperm)
        dx = float(np.abs(got[ia, 0:3] - r[ib, 0:3]).max()) if complete else float("inf")
        dv = float(np.abs(got[ia, 4:7] - r[ib, 4:7]).max()) if complete else float("inf")
        res = {"workload": f"{args.workload} N={n2}", "steps": steps, "ranks": world, "first_step_same_tree_order_as_1gpu": same_order,
               "first_step_p2p_interactions_sum": p2p_sum, "first_step_p2p_interactions_1gpu": rp2p1,
               "all_particles_present": complete, "max_abs_dx": dx, "max_abs_dv": dv,
               "m2l_interactions_sum": int(sum(p[3] for p in parts)), "m2l_interactions_1gpu": int(rst["m2l_interactions"]),
               "device_bytes_per_rank": [int(p[4]) for p in parts], "device_bytes_1gpu": int(rst["device_bytes"])}
        ok = same_order and complete and dx < 1e-5 and dv < 2e-3 and p2p_sum == rp2p1
        if not partitioned:
            res["state_identical_on_all_ranks"] = all(abs(float(x) - float(hs[0])) <= 1e-9 * abs(float(hs[0])) for x in hs)
            ok = ok and res["state_identical_on_all_ranks"]
        res["pass"] = bool(ok)
    return res


def config1_ours(args, local_rank):
    """BASELINE config 1 through the same library: uniform cube N = 4096, 10 steps (the FMM path, the reference's capacity 8)."""
    import nbody_b200
    from nbody_b200 import workloads
    n1, steps = 4096, 10
    P1 = workloads.uniform_cube(n1)
    sim = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P1, args.dt, device=local_rank, force_constant=workloads.force_constant("uniform", n1))
    for _ in range(3):
        sim.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.step()
    dt = time.perf_counter() - t0
    sim.close()
    return {"workload": "uniform cube N=4096, 10 steps (BASELINE configs[0])", "value": n1 * steps / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / steps,
            "note": "the FMM path at the library defaults (capacity 8, order 4); compare with config1 of --impl reference"}


def launches_per_step(sim, world, n_levels=None, partitioned=False):
    """Kernels of ours launched by one step() (matches the ncu launch list profiles/r02*_launches_*.csv: 121 on one GPU for a tree of
    12 levels). The level loops run to last step's depth + 1, bounded by max_depth."""
    d = int(sim.config.max_depth)
    b = min(d, int(n_levels)) if n_levels else d
    sort = 8 * 5                 # per 8-bit pass: histogram, three scan kernels, scatter
    tree = 1 + 2 * b + (1 if b < d else 0)   # init, per level (count, split), the depth check
    upsweep = 1 + b              # P2M, per level M2M
    traversal = 1 + 2 * b        # init, per round (prep, traverse)
    far = 2 + b                  # two M2L launches, per level L2L
    base = 1 + sort + 1 + tree + upsweep + traversal + far + 1   # + keys, gather, leaf kernel
    if world == 1:
        return base
    if partitioned:
        merge = 2 * (world - 1).bit_length()
        # cuts, pull, keys, merge rounds, straddle, force, gather, export pack, publish, fix-up, 2 stamps, halo (mark, 3 scan kernels,
        # fetch, translate), multipole mark + fetch, finish, rebalance, adopt
        return base + 1 + 1 + 1 + merge + 1 + 1 + 1 + 1 + 1 + 1 + 2 + 6 + 2 + 1 + 1 + 1
    multi = 2                    # partition snap, velocity half of the gather
    if int(sim.config.flags) & 64:
        multi += 2 * (world - 1).bit_length()   # NBODY_FLAG_DIST_SORT: merge rounds of two kernels
    return base + multi


def cpu_baseline(args):
    import oracle
    from nbody_b200 import workloads
    ns = args.cpu_sample
    P = workloads.plummer(ns, n_total=args.n) if args.workload == "plummer" else workloads.GENERATORS[args.workload](ns)
    have_ref = oracle.ref_lib() is not None
    t0 = time.perf_counter()
    if have_ref:
        oracle.ref_naive_run(P, 1.0, args.dt, args.cpu_steps)
    else:
        oracle.naive_step_as_written(P, 1.0, args.dt, args.cpu_steps)
    dt = time.perf_counter() - t0
    pairs = ns * (ns - 1) / 2 * args.cpu_steps / dt
    return {"value": ns * args.cpu_steps / dt, "unit": UNIT, "cores": 1, "kind": "reference" if have_ref else "port",
            "sample": f"unmodified NaiveSimulation::step() x{args.cpu_steps} on the first {ns} particles of the workload "
                      f"({pairs:.3e} pair-interactions/s; O(N^2): at N={args.n} this is {2 * pairs / max(args.n - 1, 1):.3e} particle-steps/s)",
            "host_cores_visible": os.cpu_count()}


if __name__ == "__main__":
    main()
