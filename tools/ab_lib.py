"""A/B a second build of libkbo against the in-tree one on the cfg3 step (timings from the library's own CUDA events)."""
import json, os, shutil, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
main = os.path.join(root, "kubeflow_b200", "libkbo.so")
for alt in [None] + sys.argv[1:]:
    if alt:
        shutil.copy(main, main + ".bak"); shutil.copy(alt, main)
    try:
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "profile_step.py"), "--warmup", "2", "--steps", "3"], capture_output=True, text=True).stdout
        t = json.loads(out.strip().splitlines()[-1])["timings"]
        print(alt or "baseline", {k: round(t[k], 1) for k in ("fit_ms", "cross_kernel_ms", "var_kernel_ms", "total_ms")})
    finally:
        if alt:
            shutil.move(main + ".bak", main)
