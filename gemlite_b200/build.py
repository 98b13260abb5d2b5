"""Build libgemlite_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m gemlite_b200.build [--force] [--verbose]
    python -m gemlite_b200.build --flavour NAME -DMACRO=1 [...]   # experiment build: libgemlite_b200_NAME.so
                                                                  # (load it with GLB_LIB_PATH=...; never the default)
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgemlite_b200.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")
SOURCES = ["gl_api.cu", "gl_pack.cu", "gl_actquant.cu", "gl_generic.cu", "gl_gemv.cu", "gl_gemm_tc.cu", "gl_gemm_tc2.cu", "gl_gemm_a8_tc2.cu", "gl_gemv_mx.cu", "gl_gemv_a8.cu", "gl_gemm_mid.cu", "gl_gemm_a8_mid.cu", "gl_gemm_bs.cu", "gl_gemv_packed.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-cudart", "static",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for fn in sorted(os.listdir(root)):
            if fn.endswith((".cu", ".cuh", ".h")):
                h.update(fn.encode())
                h.update(open(os.path.join(root, fn), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False, timeline=False, flavour=None, defines=()):
    """Compile every .cu to an object (parallel) and link the shared library. Returns the path.
    timeline=True builds the diagnostic flavour (decode kernel records %globaltimer phase stamps) next to it;
    flavour=NAME + defines builds an experiment library libgemlite_b200_NAME.so with extra -D macros."""
    if flavour:
        return _build(True, verbose, os.path.join(HERE, f"libgemlite_b200_{flavour}.so"),
                      os.path.join(CSRC, f"build_{flavour}"), list(defines), None)
    if timeline:
        return _build(True, verbose, os.path.join(HERE, "libgemlite_b200_timeline.so"), os.path.join(CSRC, "build_tl"),
                      ["-DGLB_GEMV_TIMELINE=1"], None)
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    return _build(force, verbose, LIB, os.path.join(CSRC, "build"), [], dig)


def _build(force, verbose, LIB, objdir, extra, dig):
    nvcc = _nvcc()
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(f"--- {src}\n{out}\n")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
            "-Xcompiler", "-fPIC", "-o", LIB, *objs, "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    if dig is not None:
        open(STAMP, "w").write(dig)
    return LIB


if __name__ == "__main__":
    _flavour = sys.argv[sys.argv.index("--flavour") + 1] if "--flavour" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, timeline="--timeline" in sys.argv,
                flavour=_flavour, defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
