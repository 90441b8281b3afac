"""Experiment harness: time one 148-job K=4096 wave under several env settings (QUILT_B200_GEO / QUILT_B200_DBG).

    python tools/exp_sweep.py "GEO=256x16" "GEO=512x8" "DBG=1" ...
Each configuration runs in its own process (the library reads the env once)."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
cfgs = sys.argv[1:] or [""]
for c in cfgs:
    env = dict(os.environ)
    extra = []
    for kv in c.split(","):
        if kv == "ALL":
            extra.append("--all-snps")
        elif kv.startswith("K="):
            extra += ["--K", kv[2:]]
        elif kv.startswith("JOBS="):
            extra += ["--jobs", kv[5:]]
        elif kv.startswith("FF="):
            extra += ["--ff", kv[3:], "--coverage", "0.5"]
        elif "=" in kv:
            k, v = kv.split("=", 1)
            env["QUILT_B200_" + k] = v
    r = subprocess.run([sys.executable, os.path.join(HERE, "prof_sweep.py"), "--K", "4096", "--jobs", "148", "--its", "8"] + extra, env=env,
                       capture_output=True, text=True)
    lines = [ln for ln in r.stderr.splitlines() if ln.startswith("{")]
    print(json.dumps({"cfg": c, "rc": r.returncode, "timing": lines[-1] if lines else r.stderr[-300:]}), flush=True)
