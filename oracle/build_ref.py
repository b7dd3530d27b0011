#!/usr/bin/env python
"""Compile the reference's OWN Cython (operator_py/cython/{bbox,bbox_self,cpu_nms}.pyx) from where
it lies under /root/reference into oracle/_ref/*.so — test infrastructure only.

Nothing from the reference is copied into the repo: sources are staged in a temp dir (the
reference tree is read-only), cpu_nms.pyx gets the 2-token NumPy-2 compatibility patch there
(np.int_t -> np.intp_t, np.int -> np.intp; cpu_nms.pyx:45,48-49 use aliases NumPy 2 removed), and
only the built extension modules land in oracle/_ref/ (git-ignored, travels to the GPU box).
The reference's own build system (setup.py with its CUDA probing) is not run.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

REF = "/root/reference/operator_py/cython"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
MODULES = ["bbox", "bbox_self", "cpu_nms"]


def main() -> int:
    if not os.path.isdir(REF):
        print("reference not present; keeping prebuilt oracle/_ref", file=sys.stderr)
        return 0
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    if all(os.path.exists(os.path.join(OUT, m + suffix)) for m in MODULES):
        return 0
    import numpy as np
    from Cython.Build import cythonize  # noqa: F401  (fail early if Cython is missing)

    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        for m in MODULES:
            src = open(os.path.join(REF, m + ".pyx")).read()
            if m == "cpu_nms":
                src = src.replace("np.int_t", "np.intp_t").replace("dtype=np.int)", "dtype=np.intp)")
            open(os.path.join(tmp, m + ".pyx"), "w").write(src)
        for m in MODULES:
            subprocess.run([sys.executable, "-m", "cython", "-3", m + ".pyx"], cwd=tmp, check=True,
                           capture_output=True)
            inc = [sysconfig.get_paths()["include"], np.get_include()]
            cmd = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-w",
                   "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]
            for i in inc:
                cmd += ["-I", i]
            cmd += ["-o", os.path.join(OUT, m + suffix), os.path.join(tmp, m + ".c")]
            subprocess.run(cmd, check=True, capture_output=True)
    print("built", [m + suffix for m in MODULES], "->", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
