"""Recipe for oracle/_ref/qeft_cuda_ref.so: the reference's OWN CUDA kernels of this path, compiled for sm_100a.

TEST INFRASTRUCTURE ONLY (the checker and the "reference kernels on the same GPU" timing in bench.py / tools).

The reference has no CPU implementation of the QuantLinear arithmetic: every forward calls its CUDA extension
(qeft/qlinear.py:248-275).  Its three kernels of the path compile from three source files with plain nvcc plus the
torch headers of this image -- no cmake, no generated code, not the reference's setup script -- so they are built here
from the sources WHERE THEY LIE under /root/reference (nothing is copied into the repo), with the reference's own
compile flags (qeft/kernel/setup_cuda.py:4-28) and `-gencode arch=compute_100a,code=sm_100a`:

    quantization_new/gemv/gemv_cuda.cu        gemv_4bit
    quantization_new/gemv/gemv_cuda_qeft.cu   gemv_4bit_qeft
    quantization_new/gemm/gemm_cuda.cu        gemm_4bit

and bound by oracle/ref_bind.cpp.

ONE DEVIATION, in the launch configuration only.  Both GEMV files launch their kernels as
`<<<num_blocks, num_threads>>>` (gemv_cuda.cu:387-..., gemv_cuda_qeft.cu:422-...) although the kernels index
`extern __shared__ uint8_t shmem[]` for the cross-warp reduction (gemv_cuda.cu:102, gemv_cuda_qeft.cu:104): no dynamic
shared memory is requested.  On B200 (sm_100a) that is a hard fault -- the unmodified build raised
cudaErrorIllegalAddress on the first gemv_4bit_qeft call (observed on the GPU box, round 1) -- so the reference's
decode path does not run on this hardware as shipped.  To still get reference-run outputs, the recipe compiles the two
GEMV files from a scratch copy in a temporary directory (deleted afterwards; never inside the repo) whose ONLY change
is the third launch parameter, `<<<num_blocks, num_threads, 2048>>>` (the reduction needs 8 warps x 2 x m x 4 floats
<= 1792 bytes).  The kernels are untouched.  gemm_cuda.cu is compiled as it is.

Outputs go to oracle/_ref/ only (git-ignored, shipped to the GPU box by gpurun like
our own .so).  The GPU box has no /root/reference: there the prebuilt module is used as it is, and tests that need it
skip when it is absent.  The module links against the libtorch of this image (same image on the box).

    python oracle/build_ref.py            # incremental; ~4 minutes from scratch (torch headers under nvcc)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("QEFT_REFERENCE_ROOT", "/root/reference")
KERNEL_DIR = os.path.join(REF, "qeft", "kernel")
SOURCES = ["quantization_new/gemv/gemv_cuda.cu", "quantization_new/gemv/gemv_cuda_qeft.cu",
           "quantization_new/gemm/gemm_cuda.cu"]
MODULE = os.path.join(OUT, "qeft_cuda_ref.so")
LAUNCH_AS_SHIPPED = "<<<num_blocks, num_threads>>>"
LAUNCH_WITH_SMEM = "<<<num_blocks, num_threads, 2048>>>"

# the reference's nvcc flags (setup_cuda.py) + the target architecture
NVCC_FLAGS = ["-O3", "-std=c++17", "-DENABLE_BF16", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
              "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
              "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__", "--expt-relaxed-constexpr",
              "--expt-extended-lambda", "--use_fast_math", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-w"]


def available() -> bool:
    return all(os.path.exists(os.path.join(KERNEL_DIR, s)) for s in SOURCES)


def _torch_dirs():
    import torch
    root = os.path.dirname(torch.__file__)
    inc = [os.path.join(root, "include"), os.path.join(root, "include", "torch", "csrc", "api", "include"),
           sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
    return inc, os.path.join(root, "lib")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose: bool = True) -> str | None:
    """Build (incrementally) and return the module path; None when the reference sources are not on this machine."""
    if not available():
        return MODULE if os.path.exists(MODULE) else None
    os.makedirs(OUT, exist_ok=True)
    inc, libdir = _torch_dirs()
    incflags = [f"-I{d}" for d in inc] + [f"-I{KERNEL_DIR}"]
    defs = ["-DTORCH_EXTENSION_NAME=qeft_cuda_ref", "-DTORCH_API_INCLUDE_EXTENSION_H"]
    jobs = []
    objs = []
    scratch = tempfile.mkdtemp(prefix="qeft_ref_build_")
    for s in SOURCES:
        src = os.path.join(KERNEL_DIR, s)
        obj = os.path.join(OUT, os.path.basename(s).replace(".cu", ".o"))
        objs.append(obj)
        if not _stale(obj, [src]):
            continue
        extra = []
        if "/gemv/" in s:
            # launch configuration only (see the module docstring): request the shared memory the kernels index
            text = open(src).read()
            n = text.count(LAUNCH_AS_SHIPPED)
            assert n > 0, f"{s}: launch pattern not found"
            patched = os.path.join(scratch, os.path.basename(s))
            with open(patched, "w") as f:
                f.write(text.replace(LAUNCH_AS_SHIPPED, LAUNCH_WITH_SMEM))
            extra = [f"-I{os.path.dirname(src)}"]        # "gemv_cuda.h", "../dequantize.cuh" resolve from the original
            src = patched
        jobs.append(["nvcc", *NVCC_FLAGS, *incflags, *extra, *defs, "-c", src, "-o", obj])
    bind_src = os.path.join(HERE, "ref_bind.cpp")
    bind_obj = os.path.join(OUT, "ref_bind.o")
    objs.append(bind_obj)
    if _stale(bind_obj, [bind_src]):
        jobs.append(["g++", "-O2", "-std=c++17", "-fPIC", *incflags, *defs, "-c", bind_src, "-o", bind_obj])

    def run(cmd):
        if verbose:
            print("[oracle/_ref]", " ".join(cmd[:1] + cmd[-3:]), flush=True)
        subprocess.run(cmd, check=True)

    try:
        with ThreadPoolExecutor(max_workers=4) as ex:
            list(ex.map(run, jobs))
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    if jobs or _stale(MODULE, objs):
        run(["g++", "-shared", "-o", MODULE, *objs, f"-L{libdir}", "-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda",
             "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart", f"-Wl,-rpath,{libdir}"])
    return MODULE


# The reference's own PYTHON of the path (qeft/qlinear.py, qeft/reorder.py), staged UNMODIFIED next to the compiled
# kernels so that the drop-in test (tests/test_reference_module_gpu.py) can run the reference's QuantLinear against this
# repo's `qeft_cuda` shim on the GPU box, where /root/reference does not exist.  baseline/_ref/ is git-ignored (the
# files never enter the history) and travels with gpurun, like oracle/_ref/.
PY_STAGE = os.path.join(os.path.dirname(HERE), "baseline", "_ref")
PY_FILES = ["qeft/__init__.py", "qeft/qlinear.py", "qeft/reorder.py", "qeft/utils/__init__.py", "qeft/utils/misc.py"]


def stage_reference_python() -> str | None:
    """Copy the reference's qlinear.py / reorder.py as they are into baseline/_ref/qeft/ (only where /root/reference
    exists); returns the directory to put on sys.path, or None when neither the reference nor a staged copy exists."""
    have_ref = all(os.path.exists(os.path.join(REF, f)) for f in PY_FILES)
    if have_ref:
        for f in PY_FILES:
            dst = os.path.join(PY_STAGE, f)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            src = os.path.join(REF, f)
            if not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read():
                shutil.copyfile(src, dst)
    return PY_STAGE if all(os.path.exists(os.path.join(PY_STAGE, f)) for f in PY_FILES) else None


def load():
    """Import oracle/_ref/qeft_cuda_ref.so (prebuilt); None if it is not there.  Never builds."""
    if not os.path.exists(MODULE):
        return None
    import importlib.util

    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("qeft_cuda_ref", MODULE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build()
    print(p if p else "reference sources not found; nothing built")
    sys.exit(0 if p else 1)
