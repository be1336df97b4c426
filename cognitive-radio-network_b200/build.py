"""Build libcrnsense.so in-tree with nvcc for sm_100a (no JIT, no torch extension machinery).

The library is a plain C-ABI shared object (include/crnsense.h); it links the CUDA runtime statically
and nothing else, so the reference's makefile can link it with `-lcrnsense`.
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libcrnsense.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
              "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]

SOURCES = ["crn_config.cpp", "crn_api.cu", "crn_synth.cu", "crn_fuse.cu", "crn_ann.cu", "crn_sense_n256.cu", "crn_sense_n512.cu",
           "crn_sense_n1024.cu", "crn_sense_n2048.cu", "crn_sense_n4096.cu", "crn_sense_n8192.cu"]
HEADERS = ["crn_internal.h", "crn_fft_regs.cuh", "crn_sense_kernel.cuh", "crn_launch.cuh",
           os.path.join(ROOT, "include", "crnsense.h")]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libcrnsense cannot be built (there is no CPU fallback)")


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS + [os.path.abspath(__file__)]:
        p = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(ARCH + NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(nvcc, src, verbose, objdir=OBJ, defs=()):
    obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
    cmd = [nvcc] + ARCH + NVCC_FLAGS + list(defs) + (["-Xptxas", "-v"] if verbose else []) + \
          ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link libcrnsense.so next to this file."""
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()
    logs = {}
    with cf.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        futs = {ex.submit(_compile, nvcc, s, verbose): s for s in SOURCES}
        objs = []
        for f in cf.as_completed(futs):
            obj, log = f.result()
            objs.append(obj)
            logs[futs[f]] = log
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + sorted(objs) + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as fh:
        fh.write(dig)
    if verbose:
        for s in SOURCES:
            sys.stderr.write("== %s\n%s" % (s, logs[s]))
    return LIB


def build_variant(name, defs, verbose=False):
    """Development aid: the same library compiled with extra -D switches into variants/libcrnsense_<name>.so
    (A/B timing of one kernel decision on the GPU box; select it with CRN_LIB=<path> at import)."""
    vdir = os.path.join(HERE, "variants")
    objdir = os.path.join(vdir, "build_" + name)
    os.makedirs(objdir, exist_ok=True)
    lib = os.path.join(vdir, "libcrnsense_%s.so" % name)
    nvcc = _nvcc()
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        res = list(ex.map(lambda s: _compile(nvcc, s, verbose, objdir, defs), SOURCES))
    r = subprocess.run([nvcc] + ARCH + ["-shared", "-o", lib] + sorted(o for o, _ in res) +
                       ["-lcudart_static", "-lpthread", "-ldl", "-lrt"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        for s, (_, log) in zip(SOURCES, res):
            sys.stderr.write("== %s\n%s" % (s, log))
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python build.py --variant NAME -DFOO -DBAR=1
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if a.startswith("-D")], "-v" in sys.argv))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
