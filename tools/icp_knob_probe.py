import os, sys, json, subprocess
for thr in ("128", "256", "512"):
    env = dict(os.environ, SFE_ICP_THREADS=thr)
    out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "2", "--cpu-sample", "2"], env=env, capture_output=True, text=True)
    d = json.loads([l for l in out.stdout.strip().splitlines() if l.startswith("{")][-1])
    print("THREADS", thr, "icp ms/step", round(d["stage_ms_per_step"]["icp"], 3), "value", round(d["value"]))
