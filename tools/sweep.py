"""Developer sweep: one process per variant (tuning knobs are read from the environment once),
three steps of the Plummer workload each, stage times of the last step on one line.
usage: python tools/sweep.py N "ENV1=a,ENV2=b;cap=32" "..."   (items separated by ';': cap=, order=, flags=)"""
import json, os, subprocess, sys
import numpy as np
sys.path.insert(0, ".")

CHILD = r'''
import sys, json, numpy as np
sys.path.insert(0, ".")
import nbody_b200
P = np.load(sys.argv[1])
cap = int(sys.argv[2]); order = int(sys.argv[3]); flags = int(sys.argv[4]); steps = int(sys.argv[5]); tau = float(sys.argv[6])
sim = nbody_b200.CudaSimulation([1, 1, 1], P, 1e-3, order=order, leaf_capacity=cap, flags=flags, low_order_tau=tau)
for _ in range(steps): sim.step()
st = sim.stats()
acc = sim.accelerations()
print("RESULT " + json.dumps({**{k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}, "acc_checksum": float(np.abs(acc).sum(dtype=np.float64))}))
sim.close()
'''

def main():
    n = int(sys.argv[1])
    from nbody_b200 import workloads
    path = f"/dev/shm/plummer_{n}.npy"
    if not os.path.exists(path):
        np.save(path, workloads.plummer(n))
    for spec in sys.argv[2:]:
        env = dict(os.environ); cap, order, flags, steps, tau = 32, 4, 0, 3, 0.13
        for item in filter(None, spec.split(";")):
            for kv in item.split(","):
                k, v = kv.split("=")
                if k == "cap": cap = int(v)
                elif k == "order": order = int(v)
                elif k == "flags": flags = int(v)
                elif k == "steps": steps = int(v)
                elif k == "tau": tau = float(v)
                else: env[k] = v
        r = subprocess.run([sys.executable, "-c", CHILD, path, str(cap), str(order), str(flags), str(steps), str(tau)], env=env, capture_output=True, text=True, timeout=300)
        res = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if not res:
            print(f"[{spec}] FAILED rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}", flush=True); continue
        d = json.loads(res[0][7:])
        keys = ("ms_total", "ms_sort", "ms_tree", "ms_upsweep", "ms_traverse", "ms_m2l", "ms_l2l", "ms_leaf", "p2p_interactions", "m2l_interactions", "n_leaves", "retries", "acc_checksum")
        print(f"[{spec}] " + " ".join(f"{k}={d[k]}" for k in keys), flush=True)

if __name__ == "__main__":
    main()
