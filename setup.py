"""Build hook: the two native libraries (CUDA C-ABI for sm_100a, host reader) are compiled in-tree by
`__graft_entry__.build()` before setuptools collects `topsicle_b200/*.so` as package data.
Metadata and the `topsicle` console script live in pyproject.toml (reference: /root/reference/setup.py:17-21)."""
import os
import sys

from setuptools import setup
from setuptools.command.build_py import build_py

HERE = os.path.dirname(os.path.abspath(__file__))


class BuildNative(build_py):
    def run(self):
        sys.path.insert(0, HERE)
        import __graft_entry__
        __graft_entry__.build()
        super().run()


setup(cmdclass={"build_py": BuildNative})
