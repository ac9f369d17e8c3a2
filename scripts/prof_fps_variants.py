import os, sys, subprocess
for T in ("", "64", "128", "512", "1024"):
    env = dict(os.environ)
    if T: env["PZ_FPS_T"] = T
    out = subprocess.run([sys.executable, "scripts/prof_geometry.py"], env=env, capture_output=True, text=True).stdout.strip()
    print("T=", T or "default(256)", out)
