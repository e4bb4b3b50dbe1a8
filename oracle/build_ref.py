"""TEST INFRASTRUCTURE ONLY — compile the reference's own native NMS (MQ/libs/utils/csrc/nms_cpu.cpp) from the
sources where they lie under /root/reference into oracle/_ref/ (git-ignored; travels to the GPU box).

Nothing is copied: g++ is pointed at the reference file directly.  The resulting python extension
`nms_1d_cpu` is what `libs/utils/nms.py` imports; it is used (a) to pin oracle/softnms.c and the CUDA
soft-NMS, (b) as the `cpu_baseline.kind == "reference"` leg for the NMS part of bench.py.
"""
import os
import subprocess
import sys
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("VILCO_REFERENCE", "/root/reference")
SRC = os.path.join(REF_ROOT, "MQ", "libs", "utils", "csrc", "nms_cpu.cpp")
OUT_DIR = os.path.join(_HERE, "_ref")


def so_path():
    return os.path.join(OUT_DIR, "nms_1d_cpu" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_nms(force=False):
    """Returns the directory holding nms_1d_cpu*.so (building it if the reference source is present)."""
    out = so_path()
    if os.path.exists(out) and not force:
        return OUT_DIR
    if not os.path.exists(SRC):
        raise RuntimeError(f"reference source {SRC} not present and {out} not prebuilt")
    import torch
    from torch.utils import cpp_extension
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = []
    for p in cpp_extension.include_paths():
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-DTORCH_EXTENSION_NAME=nms_1d_cpu",
           "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           *inc, SRC, "-o", out, f"-L{libdir}", f"-Wl,-rpath,{libdir}",
           "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python"]
    subprocess.run(cmd, check=True)
    return OUT_DIR


def load_nms():
    d = build_nms()
    if d not in sys.path:
        sys.path.insert(0, d)
    import torch  # noqa: F401  (must be imported before the extension)
    import nms_1d_cpu
    return nms_1d_cpu


if __name__ == "__main__":
    print(build_nms(force="--force" in sys.argv))
