"""Build entry point in the place the reference keeps its own (models/ops/setup.py + make.sh):

    python setup.py build_ext --inplace        # == python -m uvhand_b200.build

Compiles, for sm_100a ONLY (``-gencode arch=compute_100a,code=sm_100a``; no other architecture, no CPU
source, no fallback), the torch-free kernel library ``uvhand_b200/lib/libmsda_b200.so`` with nvcc and the
``MultiScaleDeformableAttention`` extension module with the host compiler.  Unlike the reference's setup.py
(which raises when no GPU is visible, models/ops/setup.py:36-47) this cross-compiles on a GPU-less machine.
The artefacts stay in-tree next to the sources, like the reference's ``build_ext --inplace`` flow.
"""
import os
import sys

from setuptools import Command, find_packages, setup

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


class BuildNative(Command):
    description = "build the sm_100a kernel library and the torch extension in-tree"
    user_options = [("inplace", "i", "accepted for compatibility; the build is always in-tree"),
                    ("force", "f", "rebuild even if up to date")]
    boolean_options = ["inplace", "force"]

    def initialize_options(self):
        self.inplace = True
        self.force = False

    def finalize_options(self):
        pass

    def run(self):
        from uvhand_b200 import build
        build.build_all(force=bool(self.force), verbose=True)


setup(
    name="uvhand_b200",
    version="0.1.0",
    description="B200-native (sm_100a) multi-scale deformable attention: drop-in for UVHand / Deformable-DETR MSDeformAttn",
    packages=find_packages(include=["uvhand_b200", "uvhand_b200.*"]),
    package_data={"uvhand_b200": ["lib/*.so", "*.so", "csrc/*"]},
    cmdclass={"build_ext": BuildNative},
)
