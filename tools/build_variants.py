"""Build experiment variants of the CUDA library into rajepy_b200/lib/variants/<name>.so
(git-ignored, travels to the GPU box):  python tools/build_variants.py name=-DFLAG[,-DFLAG2] ...
A name of the form  name@<git-rev>  takes csrc/rjp_integrate.cu from that revision.
Select one at run time with RAJEPY_B200_LIB=<path>."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rajepy_b200 import build as b  # noqa: E402


def main():
    out_dir = os.path.join(b.LIBDIR, "variants")
    os.makedirs(out_dir, exist_ok=True)
    procs = []
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition("=")
        flags = [f for f in flags.split(",") if f]
        rev = None
        if "@" in name:
            name, rev = name.split("@")
        src_dir = b.CSRC
        if rev:
            tmp = tempfile.mkdtemp()
            src_dir = os.path.join(tmp, "rajepy_b200", "csrc")
            shutil.copytree(b.CSRC, src_dir)
            shutil.copytree(os.path.join(ROOT, "include"), os.path.join(tmp, "include"))
            blob = subprocess.run(["git", "-C", ROOT, "show",
                                   f"{rev}:rajepy_b200/csrc/rjp_integrate.cu"],
                                  capture_output=True, text=True, check=True).stdout
            with open(os.path.join(src_dir, "rjp_integrate.cu"), "wt") as f:
                f.write(blob)
        out = os.path.join(out_dir, f"{name}.so")
        cmd = [b._nvcc()] + b.NVCC_FLAGS + flags + \
            [os.path.join(src_dir, s) for s in b.SOURCES] + ["-o", out]
        log = open(os.path.join(out_dir, f"{name}.log"), "wt")
        procs.append((name, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT)))
    for name, p in procs:
        print(name, "rc", p.wait())


if __name__ == "__main__":
    main()
