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
case = tk.Case.load(tk.make_run_dir(f"/tmp/run_{cfg}", cfg))
case.build_tables(shi_window_only=True, cache_dir=os.path.join(ROOT, ".table_cache"))
for n_eng in (1, 2, 3):
    engs = [tk.Engine(case, batch=batch) for _ in range(n_eng)]
    per = nit // n_eng
    def work(i):
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
    print(f"{n_eng} engine(s), batch {batch}: {best * 1e3:.1f} ms for {per * n_eng} iterations = {per * n_eng / best:.0f} iterations/s", flush=True)
    del engs
