"""Developer tool: build libcholespy_b200_<name>.so with extra nvcc flags (A/B runs: CHB_LIB_VARIANT=<name>).

    python scripts/build_variant.py <name> [nvcc flags ...]
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholespy_b200 import build  # noqa: E402

name = sys.argv[1]
main = os.path.join(build.LIB, "libcholespy_b200.so")
keep = main + ".keep"
shutil.copy2(main, keep)
build.NVCC_FLAGS += sys.argv[2:]
build.build_library(force=True)
shutil.copy2(main, os.path.join(build.LIB, "libcholespy_b200_%s.so" % name))
os.replace(keep, main)
os.utime(os.path.join(build.LIB, "kernels.cubin"), (0, 0))  # the cubin on disk is the variant's: force a rebuild next time
print("built variant", name)
