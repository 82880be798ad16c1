import os, sys, json, subprocess
for v in ("0", "4096"):
    env = dict(os.environ, SFE_ICP_SMALL_NT=v)
    out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "2", "--cpu-sample", "2"], env=env, capture_output=True, text=True)
    d = json.loads(out.stdout.strip().splitlines()[-1])
    print("SMALL_NT", v, "icp ms/step", d["stage_ms_per_step"]["icp"], "value", d["value"], "matched", d["config"]["frames_matched_last_step"], "mean pts", d["config"]["mean_cloud_points"])
