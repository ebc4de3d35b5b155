"""Build libbn_b200.so (sm_100a only) in-tree with nvcc.  `python -m bn_b200.build [--force]`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libbn_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "550",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inc", ".py"))) + [
        os.path.join(os.path.dirname(HERE), "include", "bn_b200.h")]


def source_hash() -> str:
    """sha256 over the kernel sources: ties a committed ncu artefact (profiles/ncu_kernels.json) to the build it profiled."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".inc")):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return SO
    if not (os.path.exists(os.path.join(CSRC, "constants_fp.inc")) and os.path.exists(os.path.join(CSRC, "constants_tower.inc"))):
        subprocess.check_call([sys.executable, os.path.join(CSRC, "gen_constants.py")])
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, os.path.join(CSRC, "kernels.cu")]
    subprocess.check_call(cmd, cwd=HERE)
    return SO


def build_variant(name: str, defines, verbose: bool = False) -> str:
    """A/B build: bn_b200/libv_<name>.so with extra -D flags (select it at run time with BN_B200_SO=<path>)."""
    out = os.path.join(HERE, "libv_%s.so" % name)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, os.path.join(CSRC, "kernels.cu")]
    subprocess.check_call(cmd, cwd=HERE)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if a.startswith("-D")], verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
