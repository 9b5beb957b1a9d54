"""Build libzfp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

One object per translation unit, compiled in parallel; the kernel instances are spread over
inst_{enc,dec}_<type>.cu so a full build takes a couple of minutes instead of ten.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build", "default")
LIB_DIR = os.path.join(HERE, "lib")
# developer knobs for A/B experiments: an alternative library name and extra nvcc flags (-D...)
LIB = os.path.join(LIB_DIR, os.environ.get("ZFP_B200_LIB_NAME", "libzfp_b200.so"))
EXTRA = os.environ.get("ZFP_B200_EXTRA_FLAGS", "").split()
if EXTRA:
    OBJ_DIR = os.path.join(HERE, "build", "variant_" + os.environ.get("ZFP_B200_LIB_NAME", "x"))
INCLUDE = os.path.join(HERE, "..", "include")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17"] + ARCH + ["-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def _headers():
    return (glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
            glob.glob(os.path.join(INCLUDE, "*.h")) + [os.path.abspath(__file__)])


def _nvcc():
    cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")]
    # the image exports CC/CXX=/opt/gcc wrappers; nvcc wants the distro host compiler
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    return cmd


def _compile(src, obj, verbose):
    cmd = _nvcc() + NVCC_FLAGS + EXTRA + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stderr, " ".join(cmd)


def build_library(force=False, verbose=False, jobs=None):
    """Compile the CUDA backend + host API into zfp_b200/lib/libzfp_b200.so; returns its path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    newest_header = max(os.path.getmtime(h) for h in _headers())
    # an up-to-date library is enough (the object files do not travel to the GPU box)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max([newest_header] + [os.path.getmtime(f) for f in _sources()]):
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    todo, objs = [], []
    for src in _sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            todo.append((src, obj))
    log = []
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=jobs or min(len(todo), os.cpu_count() or 4)) as ex:
            for src, rc, err, cmd in ex.map(lambda so: _compile(so[0], so[1], verbose), todo):
                log.append("### %s\n%s" % (os.path.basename(src), err))
                if rc != 0:
                    raise RuntimeError("nvcc failed:\n%s\n%s" % (cmd, err[-8000:]))
    if todo or not os.path.exists(LIB):
        cmd = _nvcc() + ARCH + ["-shared", "-Xlinker", "-Bsymbolic", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-8000:]))
    if verbose:
        sys.stderr.write("\n".join(log) + "\n")
    return LIB


if __name__ == "__main__":
    print(build_library(force="-f" in sys.argv, verbose="-v" in sys.argv))
