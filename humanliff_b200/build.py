"""Build recipe for ``libhumanliff_b200.so`` (sm_100a only, in-tree).

``python -m humanliff_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without
a GPU; the resulting ``.so`` lives next to this file so that it travels with the repo snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# experiment builds: HL_BUILD_TAG=<tag> HL_NVCC_EXTRA="-DX=1" -> libhumanliff_b200_<tag>.so (select with $HL_LIB)
_TAG = os.environ.get("HL_BUILD_TAG", "")
OBJ_DIR = os.path.join(HERE, "csrc", "build" + ("_" + _TAG if _TAG else ""))
LIB_PATH = os.path.join(HERE, "libhumanliff_b200" + ("_" + _TAG if _TAG else "") + ".so")

SOURCES = ["elementwise.cu", "conv_simt.cu", "conv_tc.cu", "attention.cu", "attention_tc5.cu", "render.cu", "render_tc.cu", "render_tc5.cu", "render_tc5_canon.cu", "sampler.cu", "gn_skip_tc5.cu"]

EXTRA_DEPS = {"render_tc5_canon.cu": ["render_tc5.cu"]}      # sources that #include another source

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
] + os.environ.get("HL_NVCC_EXTRA", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "humanliff_b200.h"))
    # whole-library stamp next to the .so: a snapshot that ships the library without the object directory (the GPU
    # box) must not recompile 6 translation units just to find out nothing changed
    lib_stamp = LIB_PATH + ".sha"
    lib_digest = _digest([os.path.join(CSRC, src) for src in SOURCES] + headers)
    if (not force and os.path.exists(LIB_PATH) and os.path.exists(lib_stamp)
            and open(lib_stamp).read() == lib_digest):
        return LIB_PATH
    objs = []
    relink = force or not os.path.exists(LIB_PATH)
    stale = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        stamp = op + ".sha"
        dg = _digest([sp] + headers + [os.path.join(CSRC, d) for d in EXTRA_DEPS.get(src, [])])
        fresh = (not force and os.path.exists(op) and os.path.exists(stamp)
                 and open(stamp).read() == dg)
        if not fresh:
            cmd = [nvcc] + NVCC_FLAGS + ["-c", sp, "-o", op]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), file=sys.stderr)
            stale.append((cmd, stamp, dg))
        objs.append(op)
    if stale:
        # the translation units are independent: compile them side by side (a from-scratch build is 11 nvcc runs of
        # 5 - 40 s each); the stamp of an object is written only after ITS compile succeeded
        from concurrent.futures import ThreadPoolExecutor

        def compile_one(job):
            cmd, stamp, dg = job
            subprocess.run(cmd, check=True)
            with open(stamp, "w") as f:
                f.write(dg)

        jobs = max(1, min(len(stale), int(os.environ.get("HL_BUILD_JOBS", "0")) or (os.cpu_count() or 1)))
        with ThreadPoolExecutor(max_workers=jobs) as pool:
            list(pool.map(compile_one, stale))
        relink = True
    if relink:
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    with open(lib_stamp, "w") as f:
        f.write(lib_digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
