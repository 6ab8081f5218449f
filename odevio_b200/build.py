"""In-tree build of libodevio_b200.so (nvcc, sm_100a only; cross-compiles without a GPU).

    python -m odevio_b200.build [--force] [--verbose]

The shared object lands next to this file under ``lib/`` so it travels with the repo
snapshot to the GPU box; it is git-ignored (history stays source-only).
"""

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libodevio_b200.so"
LIBPATH = os.path.join(LIBDIR, LIBNAME)
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the odevio_b200 CUDA library cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint():
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(INCLUDE, "odevio.h")]
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False, variant=None, defines=()):
    """Compile every .cu under csrc/ into one shared object.  Returns its path.
    `variant`/`defines` build a tuning variant (lib/libodevio_b200.<variant>.so) for A/B runs;
    select it at run time with ODEVIO_LIB_PATH."""
    os.makedirs(LIBDIR, exist_ok=True)
    if variant:
        return _build(os.path.join(LIBDIR, f"libodevio_b200.{variant}.so"), variant, list(defines), verbose)
    stamp = os.path.join(LIBDIR, ".fingerprint")
    fp = _fingerprint()
    if not force and os.path.exists(LIBPATH) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == fp:
                return LIBPATH
    path = _build(LIBPATH, "", [], verbose)
    with open(stamp, "w") as fh:
        fh.write(fp)
    return path


def _headers_digest():
    h = hashlib.sha256()
    for d in (CSRC, INCLUDE):
        for f in sorted(os.listdir(d)):
            if f.endswith((".h", ".cuh")):
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(f.encode() + fh.read())
    return h.hexdigest()


def _build(libpath, tag, defines, verbose):
    """Objects are compiled in parallel and reused when neither their source, any header nor the flags changed."""
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    hdr = _headers_digest()
    log = []

    def compile_one(src):
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + (f".{tag}" if tag else "") + ".o")
        cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj]
        with open(src, "rb") as fh:
            key = hashlib.sha256(fh.read() + hdr.encode() + " ".join(cmd).encode()).hexdigest()
        stamp = obj + ".key"
        if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == key:
            return obj, "$ (cached) " + obj + "\n", 0
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode == 0:
            with open(stamp, "w") as fh:
                fh.write(key)
        return obj, "$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr, r.returncode

    objs = []
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for src, (obj, text, rc) in zip(sources(), ex.map(compile_one, sources())):
            log.append(text)
            if rc != 0:
                sys.stderr.write(text)
                raise RuntimeError(f"nvcc failed on {src}")
            objs.append(obj)
    cmd = [nvcc, "-shared", "-o", libpath] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("link failed")
    with open(os.path.join(LIBDIR, f"build{'.' + tag if tag else ''}.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return libpath


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
