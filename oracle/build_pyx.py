"""Builds the reference's OWN Cython binding (python/py_corona_sim.pyx, compiled IN PLACE from /root/reference, never
copied) against this repository's observation_fit facade (3d_planetary_rt_model_b200/host/observation_fit.hpp), as the
module name the reference's setup script gives its CUDA build (py_corona_sim_gpu, setup_corona_sim.py:36-40).

It is the drop-in check of SURVEY.md 8(b): "keep every observation_fit public signature and the .pyx unchanged".
Outputs only into oracle/_ref/py_corona_sim/ (git-ignored; travels to the GPU box): the generated C++, the extension
module and a copy of the IPH table the binding expects next to the module (py_corona_sim.pyx:28-31,178-201).
Needs /root/reference; on the GPU box the prebuilt module is used.   python oracle/build_pyx.py
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("B200RT_REFERENCE", "/root/reference")
PYX = os.path.join(REF, "python", "py_corona_sim.pyx")
OUT = os.path.join(HERE, "_ref", "py_corona_sim")
PKG = os.path.join(ROOT, "3d_planetary_rt_model_b200")
MODULE = "py_corona_sim_gpu"


FAST_PYX = os.path.join(PKG, "host", "py_corona_sim_b200.pyx")
FAST_OUT = os.path.join(HERE, "_ref", "py_corona_sim_fast")


def build(verbose=True, fast=False):
    """fast=False: the reference's .pyx unchanged.  fast=True: host/py_corona_sim_b200.pyx, which INCLUDES the reference's
    .pyx from the reference tree (Cython include path) and adds the buffer-protocol / nogil subclass; same module name
    (the reference's __cinit__ locates the IPH table through it), its own output directory."""
    PYX, OUT = (FAST_PYX, FAST_OUT) if fast else (globals()["PYX"], globals()["OUT"])
    if not os.path.exists(globals()["PYX"]):
        raise FileNotFoundError(globals()["PYX"])
    import numpy
    from Cython.Compiler.Main import CompilationOptions, default_options, compile as cy_compile
    os.makedirs(OUT, exist_ok=True)
    cpp = os.path.join(OUT, MODULE + ".cpp")
    so = os.path.join(OUT, MODULE + sysconfig.get_config_var("EXT_SUFFIX"))
    deps = [PYX, globals()["PYX"], os.path.join(PKG, "host", "observation_fit.hpp"), os.path.join(PKG, "host", "atmosphere.hpp"),
            os.path.join(PKG, "host", "fast_binding.hpp"), os.path.join(PKG, "libb200rt_host.so"), __file__]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return so
    try:
        git_hash = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:
        git_hash = ""
    opts = CompilationOptions(default_options, cplus=True, language_level=3, output_file=cpp,
                              include_path=[os.path.join(REF, "python")],
                              compile_time_env={"RT_FLOAT": False, "CPP_GIT_HASH": "b200rt-" + (git_hash or "unknown")})
    res = cy_compile(PYX, opts, full_module_name=MODULE)
    if res.num_errors:
        raise RuntimeError("cython failed on the reference binding")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
           "-I" + os.path.join(PKG, "host"), "-I" + numpy.get_include(), "-I" + sysconfig.get_paths()["include"],
           cpp, "-o", so, "-L" + PKG, "-lb200rt_host", "-lb200rt", "-Wl,-rpath,$ORIGIN/../../../3d_planetary_rt_model_b200", "-Wl,-rpath," + PKG]
    if verbose:
        print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    table = os.path.join(REF, "python", "quemerais_IPH_sourcefn_fsm99td12v20t80.dat")
    if os.path.exists(table):
        shutil.copy(table, OUT)
    return so


if __name__ == "__main__":
    print(build())
    print(build(fast=True))
