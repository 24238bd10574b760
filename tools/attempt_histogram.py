"""Attempts per reactor of the config-3 ensemble (2^20 reactors, Ros4, default tolerances) -> gpurun_out/attempts_config3.npy,
and a queue simulation of the integrator's lane occupancy: how many lane-attempt slots are idle because a lane waits for
its refill slot, and how many because the queue has run dry while other lanes of the block are still integrating."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def simulate(att, lanes=148*384, refill_every=3, group=128):
    """Greedy queue: a lane takes the next reactor when it is idle at a refill slot of its group."""
    import heapq
    n = att.size
    # event-driven per lane: time in attempts; refill slots at multiples of refill_every (per lane-group phase ignored)
    free = [(0, i) for i in range(lanes)]
    heapq.heapify(free)
    busy = 0
    end = np.zeros(lanes)
    for k in range(n):
        t, i = heapq.heappop(free)
        t0 = -(-t//refill_every)*refill_every          # wait for the next refill slot
        t1 = t0 + int(att[k])
        busy += int(att[k])
        end[i] = t1
        heapq.heappush(free, (t1, i))
    T = end.max()
    # a block runs until its last lane is done; the launch until the last block
    per_block = end.reshape(-1, 384).max(axis=1)
    return {"useful_lane_attempts": busy, "launch_length_attempts": float(T), "lane_slots": float(T)*lanes,
            "occupancy_launch": busy/(float(T)*lanes), "occupancy_until_block_end": busy/float((per_block*384).sum()),
            "mean_block_end": float(per_block.mean()), "max_block_end": float(per_block.max())}


if __name__ == "__main__":
    path = os.path.join(ROOT, "gpurun_out", "attempts_config3.npy")
    if len(sys.argv) > 1 and sys.argv[1] == "gpu":
        import cases, torch
        from rmt_app_b200 import engine
        B = 1 << 20
        base = cases.methanol_readme_input("N1")
        sw = cases.config3_sweep(B, 20240611)
        cm = engine.compile_model(base, method="ros4")
        dsw = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in sw.items()}
        r = engine.n1_solve_ensemble(cm, base, dsw, B, rtol=1e-3, atol=1e-6, keep_on_device=True)
        att = r.stats[3].cpu().numpy().astype(np.int32)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.save(path, att)
        print("attempts: mean %.2f min %d max %d p99 %d" % (att.mean(), att.min(), att.max(), np.percentile(att, 99)))
    else:
        att = np.load(path)
        print("attempts: mean %.2f std %.2f min %d max %d p50 %d p99 %d" % (att.mean(), att.std(), att.min(), att.max(), np.median(att), np.percentile(att, 99)))
        for re_ in (1, 3):
            print("refill every %d:" % re_, simulate(att, refill_every=re_))
        o = np.argsort(-att, kind="stable")
        print("longest first, refill every 3:", simulate(att[o], refill_every=3))
