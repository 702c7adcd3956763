"""Tuning harness: time the fused scan kernel for every experimental build of the library (tools/exp_build.py).
Each variant runs in its own process (KTA_LIB selects the .so)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sorted(glob.glob(os.path.join(ROOT, "kafka_topic_analyzer_b200", "libkta_gpu_exp_*.so")))
extra = sys.argv[1:]
for lib in libs:
    name = os.path.basename(lib)[len("libkta_gpu_exp_"):-3]
    env = dict(os.environ, KTA_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu", "--no-e2e",
                        "--no-extra", *extra], env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print("%-14s kernel %.4f ms  frac %.3f  value %.3e" % (name, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"]), flush=True)
    except Exception as e:
        print(name, "FAILED", r.stderr[-300:], flush=True)
