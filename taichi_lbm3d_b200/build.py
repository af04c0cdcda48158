"""Build liblbm3d_b200.so (hand-written sm_100a CUDA + the C ABI of include/lbm3d.h) in-tree.

    python -m taichi_lbm3d_b200.build [--force]

nvcc cross-compiles without a GPU.  The kernel translation units are compiled twice:
production arithmetic (namespace lbm_fast) and, with -DLBM_STRICT -fmad=false, the
oracle-order verification arithmetic (namespace lbm_strict).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "liblbm3d_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
                 "-Wno-deprecated-gpu-targets"]

# (object name, source, extra flags)
UNITS = [
    ("lbm_kernels_fast.o", "lbm_kernels.cu", []),
    ("lbm_kernels_strict.o", "lbm_kernels.cu", ["-DLBM_STRICT=1", "-fmad=false"]),
    ("lbm2p_kernels_fast.o", "lbm2p_kernels.cu", []),
    ("lbm2p_kernels_strict.o", "lbm2p_kernels.cu", ["-DLBM_STRICT=1", "-fmad=false"]),
    ("lbm_api.o", "lbm_api.cu", []),
    ("lbm2p_api.o", "lbm2p_api.cu", []),
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build the CUDA extension (there is no CPU fallback)")


def _sources():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "lbm3d.h"))
    deps.append(os.path.join(HERE, "..", "include", "lbm3d_2phase.h"))
    return [d for d in deps if os.path.isfile(d)]


def have_nvcc():
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


class build_lock:
    """Exclusive lock file around needs_build()/build(): the ranks of a multi-GPU job all call
    _lib.load(), and only one of them may write lib/obj/*.o and link."""

    def __enter__(self):
        import fcntl
        os.makedirs(LIBDIR, exist_ok=True)
        self._fh = open(os.path.join(LIBDIR, ".build.lock"), "w")
        fcntl.flock(self._fh, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self._fh, fcntl.LOCK_UN)
        self._fh.close()
        return False


def _source_hash():
    """content hash of every source the library is built from (mtimes do not survive the copy to
    a GPU box, contents do)"""
    import hashlib
    h = hashlib.sha256()
    for path in sorted(_sources()):
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(os.environ.get("LBM3D_NVCC_FLAGS", "").encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB):
        return True
    try:
        with open(LIB + ".srchash") as fh:
            return fh.read().strip() != _source_hash()
    except OSError:
        return True


def build(force=False, verbose=False, tag=None, flags=None):
    """Compile and link; returns the library path.  `tag` / `flags` build a tuning variant
    lib/liblbm3d_b200.<tag>.so with extra nvcc flags (selected at run time with LBM3D_LIB)."""
    global OBJDIR, LIB
    if tag:
        saved = OBJDIR, LIB
        OBJDIR, LIB = os.path.join(LIBDIR, "obj_" + tag), os.path.join(LIBDIR, "liblbm3d_b200.%s.so" % tag)
        os.environ["LBM3D_NVCC_FLAGS"] = flags or ""
        try:
            return build(force=True, verbose=verbose)
        finally:
            OBJDIR, LIB = saved
            os.environ.pop("LBM3D_NVCC_FLAGS", None)
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJDIR, exist_ok=True)
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper without libgomp; use the system g++
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    objs = []
    procs = []
    tune = os.environ.get("LBM3D_NVCC_FLAGS", "").split()      # tuning experiments only
    for obj, src, extra in UNITS:
        srcp = os.path.join(CSRC, src)
        if not os.path.exists(srcp):
            continue
        objp = os.path.join(OBJDIR, obj)
        cmd = [nvcc] + ccbin + COMMON + extra + tune + ["-c", srcp, "-o", objp]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)))
        objs.append(objp)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            sys.stderr.write(out.decode(errors="replace"))
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out.decode(errors="replace")))
    tmp = LIB + ".tmp"
    cmd = [nvcc] + ccbin + ARCH + ["-shared", "-o", tmp] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stdout.decode(errors="replace")))
    os.replace(tmp, LIB)
    with open(LIB + ".srchash", "w") as fh:
        fh.write(_source_hash())
    return LIB


if __name__ == "__main__":
    _tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else None
    _flags = sys.argv[sys.argv.index("--flags") + 1] if "--flags" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, tag=_tag, flags=_flags))
