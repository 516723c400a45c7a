#!/usr/bin/env python3
"""Turn one full-step ncu capture (profiles/run_ncu.sh) into the committed evidence of a round:
  profiles/<round>/ncu_summary_<tag>.txt|json  one line per profiled launch
  profiles/<round>/launches_<tag>.csv          the gpu__time_duration launch list of the same command
  profiles/traffic.json                        measured DRAM bytes per IMAGE and stage (read + write), used by bench.py for roofline.traffic
usage: collect.py <tag> <round-dir> <images-in-capture>"""
import csv, io, json, os, shutil, subprocess, sys

tag, rdir, nimg = sys.argv[1], sys.argv[2], int(sys.argv[3])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(root, "gpurun_out", f"prof_{tag}.ncu-rep")
os.makedirs(os.path.join(root, "profiles", rdir), exist_ok=True)
txt = subprocess.run([sys.executable, os.path.join(root, "profiles", "ncu_summary.py"), rep, "--json", os.path.join(root, "profiles", rdir, f"ncu_summary_{tag}.json")],
                     capture_output=True, text=True).stdout
open(os.path.join(root, "profiles", rdir, f"ncu_summary_{tag}.txt"), "w").write(txt)
src = os.path.join(root, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(root, "profiles", rdir, f"launches_{tag}.csv"))
rows = json.load(open(os.path.join(root, "profiles", rdir, f"ncu_summary_{tag}.json")))
stage_of = lambda k: ("pyramid" if "resize" in k or "margin" in k else "fast" if "fast" in k else "blur" if "blur" in k else "octree" if "octree" in k
                      else "assemble" if "assemble" in k else "orient_describe" if "describe" in k else "other")
# the capture window is one step long or a launch or two longer: every launch position of a step must count exactly once,
# so keep the LAST `per_step` rows (per_step = number of distinct launches of a step = rows until the first kernel+grid repeats)
keys = [(r["kernel"], r.get("grid"), round(r.get("dram_rd_MB", 0), 1)) for r in rows]
per_step = len(rows)
for n in range(1, len(rows)):
    if keys[n:] == keys[:len(rows) - n]:
        per_step = n
        break
rows = rows[-per_step:]
per_stage, us, winst, issue_w, occ_w = {}, {}, {}, {}, {}
for r in rows:
    st = stage_of(r["kernel"])
    per_stage[st] = per_stage.get(st, 0.0) + (r.get("dram_rd_MB", 0) + r.get("dram_wr_MB", 0)) * 1e6 / nimg
    us[st] = us.get(st, 0.0) + r.get("us", 0)
    winst[st] = winst.get(st, 0.0) + r.get("warp_inst", 0)
    issue_w[st] = issue_w.get(st, 0.0) + r.get("issue_pct", 0) * r.get("us", 0)   # time-weighted over the stage's launches
    occ_w[st] = occ_w.get(st, 0.0) + r.get("occ_pct", 0) * r.get("us", 0)
out = {"launches_per_step": per_step, "source": f"ncu --set full, {tag}, {nimg} KITTI-shaped images per launch set", "images": nimg, "bytes_per_image": per_stage,
       "us_per_launch_set": us, "warp_instructions_per_image": {k: v / nimg for k, v in winst.items()},
       "issue_pct": {k: issue_w[k] / us[k] for k in us if us[k]}, "achieved_occupancy_pct": {k: occ_w[k] / us[k] for k in us if us[k]}}
json.dump(out, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
