"""One small process that touches every kernel family, for ONE compute-sanitizer invocation (scripts/sanitize.sh): both
networks forward + backward in train and frozen-BatchNorm mode (the window conv variant forced on with QEB_WIN=2), the fused
head / jittered forward, CTC, decode, Levenshtein, selection, jitter, crop/pad, Adam."""
import os, sys, runpy
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for which in ("crnn", "unet"):
    sys.argv = ["dev_san.py", which]
    runpy.run_path(os.path.join(root, "scripts", "dev_san.py"), run_name="__main__")
runpy.run_path(os.path.join(root, "scripts", "dev_san_aux.py"), run_name="__main__")
print("all done")
