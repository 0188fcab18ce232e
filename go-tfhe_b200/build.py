"""Builds the native libraries in-tree (go-tfhe_b200/lib/):
  libtfhe_b200.so         CUDA engine + C ABI, nvcc for sm_100a only
  libtfhe_b200_client.so  host-side client helpers, g++
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(HERE, "lib")
CSRC = os.path.join(HERE, "csrc")
ENGINE = os.path.join(LIBDIR, "libtfhe_b200.so")
CLIENT = os.path.join(LIBDIR, "libtfhe_b200_client.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# translation units of the engine: (source, extra flags).  The throughput blind-rotation kernel gets its own ptxas
# register-allocation level (see blind_rotate_throughput.cu); everything else is compiled with the defaults.
ENGINE_UNITS = [("tfhe_b200.cu", []), ("blind_rotate_throughput.cu", ["-Xptxas", "--register-usage-level=7"])]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    inc = os.path.join(os.path.dirname(HERE), "include")
    eng_src = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
              [os.path.join(inc, "tfhe_b200.h")]
    if force or _stale(ENGINE, eng_src):
        objs, procs, log = [], [], ""
        for src, extra in ENGINE_UNITS:     # compiled side by side
            obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
            objs.append(obj)
            procs.append(subprocess.Popen([_nvcc()] + NVCC_FLAGS + extra + ["-c", "-o", obj, os.path.join(CSRC, src)],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        failed = False
        for p in procs:
            log += p.communicate()[0]
            failed |= p.returncode != 0
        if not failed:
            res = subprocess.run([_nvcc(), "-shared", "-o", ENGINE] + objs, capture_output=True, text=True)
            log += res.stdout + res.stderr
            failed = res.returncode != 0
        for obj in objs:
            if os.path.exists(obj):
                os.remove(obj)
        if verbose or failed:
            print(log)
        if failed:
            raise RuntimeError("nvcc failed")
        with open(os.path.join(LIBDIR, "ptxas_info.txt"), "w") as f:
            f.write(log)
    cl_src = [os.path.join(CSRC, "client.cpp"), os.path.join(CSRC, "chacha.h"), os.path.join(inc, "tfhe_b200.h"),
              os.path.join(inc, "tfhe_b200_client.h")]
    if force or _stale(CLIENT, cl_src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", CLIENT,
                               os.path.join(CSRC, "client.cpp")])
    return ENGINE, CLIENT


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose=True))
