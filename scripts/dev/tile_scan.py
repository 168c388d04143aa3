import sys, os, subprocess
here = os.path.dirname(os.path.abspath(__file__))
for tile in (0, 1, 2):
    env = dict(os.environ, PMC_TILE=str(tile), PMC_NOMAIN="1")
    code = "import sys; sys.path.insert(0, %r); from quick_time import run; run(2**24, 0.70, 20)" % here
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("tile", tile, out.stdout.strip().split("\n")[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
