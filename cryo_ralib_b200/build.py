"""In-tree build of libcryo_ralib.so (sm_100a only) with nvcc.

The library is the C-ABI drop-in for the reference's ``cuda/gpu_aln_pack.so``
(built there by ``install.sh:25``).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libcryo_ralib.so")
SOURCES = ["cra_api.cu", "cra_polar.cu", "cra_polar_grp.cu", "cra_ccf.cu", "cra_ccf_mma.cu", "cra_ccf_tm.cu", "cra_ccf_um.cu", "cra_rotsum.cu", "cra_refavg.cu", "cra_refupdate.cu", "cra_compat.cu", "cra_host.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
EXTRA = os.environ.get("CRA_NVCC_EXTRA", "").split()
FLAGS = EXTRA + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-Xcompiler", "-fopenmp", "-ccbin", "/usr/bin/g++"]


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "cryo_ralib.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, defines=()):
    """variant: build a second copy of the library with extra -D flags into experiments/bin/libcryo_ralib_<variant>.so
    (A/B measurements: CRA_LIBRARY=<that path> selects it, lib.py); the product library is not touched."""
    so, objdir, flags = SO, os.path.join(HERE, "build"), FLAGS
    if variant:
        out = os.path.join(HERE, "..", "experiments", "bin")
        os.makedirs(out, exist_ok=True)
        so = os.path.abspath(os.path.join(out, "libcryo_ralib_%s.so" % variant))
        objdir = os.path.join(HERE, "build", "variant_" + variant)
        flags = list(defines) + FLAGS
    elif not force and not _stale():
        return SO
    objs = []
    procs = []
    os.makedirs(objdir, exist_ok=True)
    for s in SOURCES:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % s)
    cmd = [NVCC, "-shared", "-o", so] + objs + ["-ccbin", "/usr/bin/g++", "-lcudart", "-lgomp"]
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    if "--variant" in sys.argv:          # python -m cryo_ralib_b200.build --variant NAME -DFLAG ...
        i = sys.argv.index("--variant")
        print(build(variant=sys.argv[i + 1], defines=[a for a in sys.argv[i + 2:] if a.startswith("-D")]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
