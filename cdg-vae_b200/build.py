"""Build recipe for libcdgvae_sm100.so: plain nvcc, sm_100a only, in-tree output so that the
library travels with a repo snapshot.  No CPU fallback is built."""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcdgvae_sm100.so")
LIB_EXP = os.path.join(HERE, "libcdgvae_sm100_exp.so")      # -DCDG_EXPERIMENTS: A/B switches read from the environment (development)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include")]


class NoCompiler(RuntimeError):
    """nvcc is absent: the only build failure after which a prebuilt library may be used."""


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise NoCompiler("nvcc not found: libcdgvae_sm100.so cannot be built")
    return nvcc


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale(experiments=False):
    lib = LIB_EXP if experiments else LIB
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(ROOT, "include", "cdgvae.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, experiments=False):
    """Compile every csrc/*.cu for sm_100a and link libcdgvae_sm100.so next to this file.

    Safe under torchrun: ranks that find the library stale at the same time serialise on a file lock, the first one
    builds (objects in a private directory, library linked under a temporary name and renamed into place atomically),
    the others re-check and find it fresh.  (Eight ranks once rebuilt it concurrently and one of them loaded a
    half-linked file: `undefined symbol`.)"""
    import fcntl
    lib = LIB_EXP if experiments else LIB
    if not force and not stale(experiments):
        return lib
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not stale(experiments):
                return lib
            return _build_locked(verbose, experiments)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose, experiments=False):
    nvcc = _nvcc()
    LIB = LIB_EXP if experiments else globals()["LIB"]
    flags = NVCC_FLAGS + (["-DCDG_EXPERIMENTS"] if experiments else [])
    objdir = os.path.join(HERE, "build", f"obj.{os.getpid()}")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + flags + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    tmp = f"{LIB}.tmp.{os.getpid()}"
    cmd = [nvcc, "-shared", "-o", tmp] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB)
    shutil.rmtree(objdir, ignore_errors=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
