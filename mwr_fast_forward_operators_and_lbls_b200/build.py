"""In-tree build of libmwrt.so (hand-written CUDA for sm_100a behind the C ABI of include/mwrt.h).

`python -m mwr_fast_forward_operators_and_lbls_b200.build` or `__graft_entry__.build()`.
nvcc cross-compiles without a GPU; the .so stays in the package directory (git-ignored, but it
travels to the GPU box with the gpurun snapshot).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("MWRT_LIB_OUT") or os.path.join(HERE, "libmwrt.so")  # override: experiments only
SOURCES = ["mwrt_api.cu", "mwrt_absorption.cu", "mwrt_rte.cu", "mwrt_rte_fanout.cu", "mwrt_post.cu",
           "mwrt_peak.cu", "mwrt_push.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--fmad=true"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmwrt.so cannot be built (there is no CPU fallback)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inc", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "mwrt.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link libmwrt.so.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = [os.path.join(CSRC, src.replace(".cu", ".o")) for src in SOURCES]
    cmds = [[nvcc] + NVCC_FLAGS + os.environ.get("MWRT_NVCC_EXTRA", "").split()
            + ["-c", os.path.join(CSRC, src), "-o", obj] for src, obj in zip(SOURCES, objs)]
    # the translation units are independent: compile them side by side (kernel 3's template family
    # alone takes ~2 minutes of ptxas)
    procs = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for c in cmds]
    log = []
    failed = None
    for c, pr in zip(cmds, procs):
        out, _ = pr.communicate()
        log.append("$ " + " ".join(c) + "\n" + out)
        if pr.returncode != 0 and failed is None:
            failed = log[-1]
    if failed is not None:
        raise RuntimeError("nvcc failed:\n" + failed)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                  "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + log[-1])
    with open(os.path.join(CSRC, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
