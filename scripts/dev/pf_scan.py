import sys, os, subprocess
here = os.path.dirname(os.path.abspath(__file__))
for env_over in ({}, {"PMC_DBG_SKIP": "32"}, {"PMC_PREFETCH": "0"}, {"PMC_PREFETCH": "148"}, {"PMC_PREFETCH": "444"},
                 {"PMC_PREFETCH": "592"}, {"PMC_PREFETCH": "888"}, {"PMC_PREFETCH": "1184"}):
    env = dict(os.environ, PMC_NOMAIN="1", **env_over)
    code = "import sys; sys.path.insert(0, %r); from quick_time import run; run(2**24, 0.70, 100)" % here
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(env_over, out.stdout.strip().split("\n")[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
