"""Build a kernel-variant copy of libhhd.so for A/B runs on the GPU box:
    python tools/build_variant.py NAME -DHHD_NEURON_MINBLOCKS_EULER=5 ...   ->  csrc/libhhd_NAME.so   (use with HHD_LIB=...)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
srcs = sorted(os.path.join(ge.CSRC, f) for f in os.listdir(ge.CSRC) if f.endswith(".cu"))
out = os.path.join(ge.CSRC, "libhhd_%s.so" % name)
cmd = [ge.NVCC] + ge.NVCC_FLAGS + flags + ["-o", out] + srcs
nccl = ge._nccl_paths()
if nccl:
    cmd += ["-DHHD_WITH_NCCL", "-I", nccl[0], "-L", nccl[1], "-l:libnccl.so.2", "-Xlinker", "-rpath=" + nccl[1]]
subprocess.check_call(cmd, cwd=ge.CSRC)
print(out)
