"""ncu raw page (csv) of EVERY launch of one batch + the run statistics -> per-kernel-class figures for bench.py's roofline.
Usage: ncu_to_json.py RAW.csv STATS.json OUT.json

Per class: launches, summed duration, DRAM bytes (read + write), warp instructions, and -- weighted by duration -- L2 hit rate,
FP64 pipe utilisation, issue-slot utilisation, threads per instruction; collisions of the class from the device counters of the
same run, hence warp instructions per collision and DRAM bytes per launch for exactly the launches whose algorithmic bytes
bench.py counts."""
import csv
import json
import re
import sys

raw, stats, out = sys.argv[1:4]
rows = list(csv.reader(open(raw)))
h, units = rows[0], rows[1]
ix = {k: i for i, k in enumerate(h)}
st = json.load(open(stats))
ev, cold, warm = st["events"], st["cold_events"], st["warm_events"]


def unit_scale(col):
    u = units[ix[col]].lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(u, 1.0)


def val(r, col):
    try:
        return float(r[ix[col]].replace(",", "")) * unit_scale(col)
    except (ValueError, KeyError):
        return 0.0


def classify(name, seen):
    m = re.search(r"k_hot<\(?(?:int\))?(\d)", name)
    if m:
        return "k_wave<electron,hot>" if m.group(1) == "0" else "k_wave<vbhole,hot>"
    m = re.search(r"k_wave<\(?(?:int\))?(\d), \(?(?:bool\))?(\d)", name)
    if m:
        sp, is_cold = int(m.group(1)), int(m.group(2))
        if sp == 2:
            return "k_wave<corehole>"
        if sp == 3:
            return "k_wave<photon>"
        base = "electron" if sp == 0 else "vbhole"
        return f"k_wave<{base},coldwarm>"          # cold and warm launches share the instantiation: split below by launch order
    for k in ("k_shi_emit", "k_shi", "k_ion_emit", "k_snapshot", "k_fold", "k_iter_prefix", "k_gen_reset", "k_ion_reset", "k_axpy", "k_companion"):
        if k + "(" in name or k + "<" in name or name.strip().endswith(k):
            return k
    return "other"


launches = []
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    launches.append({"name": name, "cls": classify(name, None), "ms": val(r, "gpu__time_duration.sum"),
                     "dram": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"), "inst": val(r, "smsp__inst_executed.sum"),
                     "l2": val(r, "lts__t_sector_hit_rate.pct"), "l1": val(r, "l1tex__t_sector_hit_rate.pct"),
                     "fp64": val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                     "issue": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "lanes": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                     "no_inst": val(r, "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
                     "barrier": val(r, "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio")})
# the LAST launch of each cold/warm instantiation is the cold one (launched once, after the hot cascade has died out)
for base in ("electron", "vbhole"):
    idx = [i for i, l in enumerate(launches) if l["cls"] == f"k_wave<{base},coldwarm>"]
    for i in idx[:-1]:
        launches[i]["cls"] = f"k_wave<{base},warm>"
    if idx:
        launches[idx[-1]]["cls"] = f"k_wave<{base},cold>"
coll = {
    "k_wave<electron,hot>": ev["el_inelastic"] + ev["el_elastic"] - cold["electron"] - warm["electron"],
    "k_wave<vbhole,hot>": ev["vbh_inelastic"] + ev["vbh_elastic"] - cold["vbhole"] - warm["vbhole"],
    "k_wave<electron,cold>": cold["electron"], "k_wave<vbhole,cold>": cold["vbhole"],
    "k_wave<electron,warm>": warm["electron"], "k_wave<vbhole,warm>": warm["vbhole"],
    "k_wave<corehole>": ev["auger"] + ev["radiative"] + ev["auger_frozen"], "k_wave<photon>": ev["photon"], "k_shi": ev["shi"],
}
res = {}
for l in launches:
    c = res.setdefault(l["cls"], {"launches": 0, "ms": 0.0, "dram_bytes": 0.0, "warp_inst": 0.0, "_w": {k: 0.0 for k in ("l2", "l1", "fp64", "issue", "lanes", "no_inst", "barrier")}})
    c["launches"] += 1; c["ms"] += l["ms"]; c["dram_bytes"] += l["dram"]; c["warp_inst"] += l["inst"]
    for k in c["_w"]:
        c["_w"][k] += l[k] * l["ms"]
for cls, c in res.items():
    w = c.pop("_w")
    t = max(c["ms"], 1e-12)
    c.update({"l2_hit_pct": w["l2"] / t, "l1_hit_pct": w["l1"] / t, "fp64_pipe_pct": w["fp64"] / t, "issue_slots_pct": w["issue"] / t,
              "threads_per_inst": w["lanes"] / t, "stall_no_instruction": w["no_inst"] / t, "stall_barrier": w["barrier"] / t})
    c["dram_bytes_per_launch"] = c["dram_bytes"] / c["launches"]
    if coll.get(cls):
        c["collisions"] = coll[cls]
        c["warp_inst_per_collision"] = c["warp_inst"] / coll[cls]
json.dump({"_comment": "ncu --set full --clock-control none of EVERY launch of one batch (scripts/ncu_round2.sh); times are cold-cache and "
                       "serialised (shares, not absolutes); collisions = device counters of the same run", "run": st, "classes": res},
          open(out, "w"), indent=1)
print("wrote", out, {k: (v["launches"], round(v["ms"], 3)) for k, v in res.items()})
