"""Builds libdiffsol_b200.so (the C-ABI library of include/diffsol_b200.h) in-tree with nvcc for sm_100a.

One translation unit per built-in equation set (csrc/dsb_inst.cu with -DDSB_INST=<id>) plus the
host/C-ABI unit, compiled in parallel and linked with the static CUDA runtime so that the library
loads through ctypes without torch.  `--fmad=false` and `-ffp-contract=off` are part of the
numerical contract: no fused multiply-add the source does not spell out, on either side.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# DSB_LIB_TAG / DSB_NVCC_EXTRA: build and load an alternative variant side by side (kernel tuning experiments)
_TAG = os.environ.get("DSB_LIB_TAG", "")
OUT_DIR = os.path.join(HERE, "_lib" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(OUT_DIR, "libdiffsol_b200.so")
N_MODELS = 23

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-diag-suppress", "128",
] + os.environ.get("DSB_NVCC_EXTRA", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _sources_digest():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode()); h.update(f.read())
    with open(os.path.join(HERE, "..", "include", "diffsol_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def have_nvcc():
    import shutil
    cand = _nvcc()
    return os.path.exists(cand) if os.path.isabs(cand) else shutil.which(cand) is not None


def is_current():
    """True when the library on disk was built from the sources as they are now."""
    stamp = os.path.join(OUT_DIR, "build.sha256")
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == _sources_digest()


def _run(cmd):
    r = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build(force=False, verbose=False):
    """Compile if the sources changed since the last build; returns the library path."""
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "build.sha256")
    digest = _sources_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    jobs = [("dsb_capi.o", [nvcc] + NVCC_FLAGS + ["-c", "dsb_capi.cu", "-o", os.path.join(OUT_DIR, "dsb_capi.o")])]
    for k in range(N_MODELS):
        obj = "dsb_inst_%d.o" % k
        jobs.append((obj, [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) +
                     ["-DDSB_INST=%d" % k, "-c", "dsb_inst.cu", "-o", os.path.join(OUT_DIR, obj)]))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        outs = list(ex.map(lambda j: _run(j[1]), jobs))
    if verbose:
        for (name, _), out in zip(jobs, outs):
            print("==", name); print(out)
    objs = [os.path.join(OUT_DIR, name) for name, _ in jobs]
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
