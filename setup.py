"""Build/packaging — same make-driven pattern as the reference's setup.py (setup.py:26-71: a custom
build_ext shells out to `make`), but sm_100a only and with no TensorFlow requirement: the C-ABI
library libfa_b200.so is always built; the TensorFlow op shim only where TensorFlow is importable.

`make` runs BEFORE build_py collects package data (so a clean `pip install .` ships the library it just
built), and build_ext additionally copies the built library into build_lib like the reference's build_ext
does with its flash_attention.so."""
import os
import shutil
import subprocess

from setuptools import Extension, setup
from setuptools.command.build_ext import build_ext
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = "tf_flash_attention_b200"
LIB = "libfa_b200.so"


def run_make():
    csrc = os.path.join(ROOT, PKG, "csrc")
    subprocess.check_call(["make", "-j", str(min(8, os.cpu_count() or 1))], cwd=csrc)
    subprocess.check_call(["make", "tf_shim"], cwd=csrc)


class MakeFirstBuildPy(build_py):
    def run(self):
        run_make()
        super().run()


class MakeBuild(build_ext):
    def run(self):
        run_make()
        if not self.inplace:
            dst = os.path.join(self.build_lib, PKG)
            os.makedirs(dst, exist_ok=True)
            shutil.copy2(os.path.join(ROOT, PKG, LIB), os.path.join(dst, LIB))


setup(
    name=PKG,
    version="0.2.0",
    description="B200-native (sm_100a, tcgen05/TMEM/TMA) drop-in engine for tf_flash_attention",
    packages=[PKG, PKG + ".tests"],
    package_data={PKG: [LIB]},
    ext_modules=[Extension(PKG + "._native", sources=[])],
    cmdclass={"build_py": MakeFirstBuildPy, "build_ext": MakeBuild},
)
