"""Build libqeb_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

One translation unit per .cu under csrc/, compiled in parallel, linked into a single shared library with the
C ABI declared in include/qeb.h. Objects are cached on source mtime + flags.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libqeb_sm100.so")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr", "-Xptxas", "-v",
    "-I", os.path.join(os.path.dirname(HERE), "include"),
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(src):
    h = hashlib.sha1()
    h.update(" ".join(FLAGS).encode())
    for f in [src] + sorted(os.path.join(CSRC, x) for x in os.listdir(CSRC) if x.endswith((".cuh", ".h"))):
        h.update(f.encode())
        h.update(str(os.stat(f).st_mtime_ns).encode())
    inc = os.path.join(os.path.dirname(HERE), "include", "qeb.h")
    if os.path.exists(inc):
        h.update(str(os.stat(inc).st_mtime_ns).encode())
    return h.hexdigest()


def _compile(name, verbose):
    src = os.path.join(CSRC, name)
    obj = os.path.join(BUILD, name[:-3] + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp(src)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, False, ""
    cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    with open(obj + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    return obj, True, r.stderr


def build(verbose=False, force=False):
    os.makedirs(BUILD, exist_ok=True)
    if force:
        for f in os.listdir(BUILD):
            os.remove(os.path.join(BUILD, f))
    names = _sources()
    objs, rebuilt = [], False
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(names))) as ex:
        for obj, did, log in ex.map(lambda n: _compile(n, verbose), names):
            objs.append(obj)
            rebuilt |= did
            if verbose and did:
                print(log)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs+ ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
