"""Two engine handles on one GPU, each running half of the iterations from its own host thread: the latency-bound hot
cascade of one overlaps the issue-bound cold phase of the other.  Usage: dual.py CFG NIT BATCH"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import trekis3_b200 as tk

cfg, nit, batch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
delays_ms = [float(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0.0]
case = tk.Case.load(tk.make_run_dir(f"/tmp/run_{cfg}", cfg))
case.build_tables(shi_window_only=True, cache_dir=os.path.join(ROOT, ".table_cache"))
for n_eng, delay in [(1, 0.0)] + [(n, d) for n in (2, 3, 4) for d in delays_ms]:
    engs = [tk.Engine(case, batch=batch) for _ in range(n_eng)]
    per = nit // n_eng
    def work(i):
        if delay > 0 and i:
            t_end = time.perf_counter() + i * delay * 1e-3
            while time.perf_counter() < t_end:
                pass
        engs[i].run_device(i * per, (i + 1) * per)
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize()
        t = time.perf_counter()
        th = [threading.Thread(target=work, args=(i,)) for i in range(n_eng)]
        [x.start() for x in th]; [x.join() for x in th]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        if rep:
            best = min(best, dt)
    print(f"{n_eng} engine(s), start offset {delay} ms, batch {batch}: {best * 1e3:.1f} ms for {per * n_eng} iterations = {per * n_eng / best:.0f} iterations/s", flush=True)
    del engs
