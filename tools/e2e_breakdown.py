"""Where does the end-to-end replan (bench.py replan_e2e) spend its wall time?  Each sub-call timed with a sync after it
(attribution), plus the host-side time of each call without waiting."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402


def main():
    torch.cuda.set_device(0)
    P = bench.GpuPlanner(0, 1024, 64, overlap=True)
    m, ff, opt = P.m, P.ff, P.opt
    for _ in range(3):
        P.replan_e2e()
    torch.cuda.synchronize()
    T = {}

    def timed(name, fn, sync=True):
        t0 = time.perf_counter()
        r = fn()
        t1 = time.perf_counter()
        if sync:
            torch.cuda.synchronize()
        t2 = time.perf_counter()
        a = T.setdefault(name, [0.0, 0.0])
        a[0] += t1 - t0
        a[1] += t2 - t0
        return r

    reps = 20
    for _ in range(reps):
        timed("upload", m.upload)
        timed("frontier_begin", P._frontier_begin)
        timed("esdf_update", m.updateESDF3d)
        timed("esdf_download", lambda: m.download(wait=True))
        timed("optimizeBatch", lambda: opt.optimizeBatch(P.x_host, P.tcs, 20, P.mask, P.evals, xtol_rel=0.0))
        timed("frontier_end", ff.search_box_end)
    for k, (h, s) in T.items():
        print("%-16s host %7.1f us   host+sync %7.1f us" % (k, 1e6 * h / reps, 1e6 * s / reps))
    # the real (overlapped) sequence of replan_e2e with host timestamps between the calls, no extra syncs
    names = ["upload_async", "frontier_begin", "esdf_update", "optimizeBatchBegin", "download_async", "frontier_end",
             "optimizeBatchEnd", "synchronize"]
    acc = np.zeros(len(names))
    for _ in range(reps):
        torch.cuda.synchronize()
        ts = [time.perf_counter()]
        m.upload(wait=False); ts.append(time.perf_counter())
        P._frontier_begin(); ts.append(time.perf_counter())
        m.updateESDF3d(); ts.append(time.perf_counter())
        opt.optimizeBatchBegin(P.x_host, P.tcs, 20, P.mask, P.evals, xtol_rel=0.0, exact_evals=True); ts.append(time.perf_counter())
        m.download(wait=False); ts.append(time.perf_counter())
        ff.search_box_end(); ts.append(time.perf_counter())
        opt.optimizeBatchEnd(out=P.opt_out); ts.append(time.perf_counter())
        m.synchronize(); ts.append(time.perf_counter())
        acc += np.diff(ts)
    print("overlapped sequence, host time per call (us):", {k: round(1e6 * v / reps, 1) for k, v in zip(names, acc)},
          "total", round(1e6 * acc.sum() / reps, 1))
    P.replan_e2e()
    torch.cuda.synchronize()
    print("device timeline of one replan_e2e (us after the upload starts):",
          {k: (round(1e3 * a, 1), round(1e3 * b, 1)) for k, (a, b) in m.last_timeline().items()})
    for order in (True, False):
        P.solver_first = order
        P.replan_e2e()
        torch.cuda.synchronize()
        print("solver_first=%s device timeline (us):" % order,
              {k: (round(1e3 * a, 1), round(1e3 * b, 1)) for k, (a, b) in m.last_timeline().items()})
        t0 = time.perf_counter()
        for _ in range(reps):
            P.replan_e2e()
        torch.cuda.synchronize()
        print("solver_first=%s replan_e2e wall %.1f us" % (order, 1e6 * (time.perf_counter() - t0) / reps))
        t0 = time.perf_counter()
        for _ in range(reps):
            P.replan_resident()
        torch.cuda.synchronize()
        print("solver_first=%s replan_resident wall %.1f us" % (order, 1e6 * (time.perf_counter() - t0) / reps))


if __name__ == "__main__":
    main()
