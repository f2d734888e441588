"""Build libs2st_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python speech-to-speech-translation_b200/build.py [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libs2st_b200.so")
SOURCES = ["api.cu", "gl_kernels.cu", "frontend_kernels.cu", "mel_tc.cu", "transform_kernels.cu", "dtw_kernels.cu", "rng_kernels.cu"]
HEADERS = ["common.cuh", "plan.h", "fft32.cuh", "frame_fft.cuh", "gl_frames.cuh", os.path.join("..", "..", "include", "s2st_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "--shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cand = os.path.join(cuda_home, "bin", "nvcc")
    return cand if os.path.isfile(cand) else "nvcc"


def is_stale():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=OUT, defines=()):
    """out / defines: experiment builds (tools/), e.g. build(out="/tmp/x.so", defines=["S2ST_GL_UNROLL_H=2"])."""
    if not force and out == OUT and not is_stale():
        return OUT
    cmd = ([_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] +
           ["-o", out] + SOURCES)
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libs2st_b200.so")
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    print(build(force="--force" in sys.argv or bool(defs), verbose="-v" in sys.argv,
                out=os.path.abspath(outs[0]) if outs else OUT, defines=defs))
