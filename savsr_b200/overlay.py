"""Swap savsr_b200's SAVSR into an unmodified checkout of the reference.

    python -m savsr_b200.overlay /path/to/SAVSR lbasicsr/test.py -opt options/test/SAVSR/test_SAVSR_Vid4_asBI.yml

The reference's ``lbasicsr/archs/__init__.py:13-16`` imports every ``*_arch.py`` next to itself by module
name, and its registry asserts uniqueness (``utils/registry.py:15``), so the new class has to *replace*
``lbasicsr.archs.savsr_arch``.  ``install()`` does that with a meta-path finder that serves this repo's
``savsr_b200/archs/savsr_arch.py`` under the reference's module name; everything else (YAML options,
model wrappers, datasets, test.py) runs unchanged.
"""
from __future__ import annotations

import importlib.abc
import importlib.util
import os
import runpy
import sys
import types

ARCH_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "archs", "savsr_arch.py")
TARGET = "lbasicsr.archs.savsr_arch"


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        if name == TARGET:
            return importlib.util.spec_from_file_location(name, ARCH_FILE)
        return None


def install(reference_root: str) -> None:
    """Make `import lbasicsr` resolve to the reference tree with SAVSR served from savsr_b200."""
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (repo_root, reference_root):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "lbasicsr.version" not in sys.modules:
        # setup.py:59-73 generates lbasicsr/version.py; a pristine checkout lacks it (lbasicsr/__init__.py:10)
        if not os.path.exists(os.path.join(reference_root, "lbasicsr", "version.py")):
            v = types.ModuleType("lbasicsr.version")
            v.__version__, v.__gitsha__, v.version_info = "0.1.1", "unknown", (0, 1, 1)
            sys.modules["lbasicsr.version"] = v
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2:
        raise SystemExit("usage: python -m savsr_b200.overlay <reference_root> <script relative to it> [script args...]")
    root, script = os.path.abspath(argv[0]), argv[1]
    install(root)
    sys.argv = [os.path.join(root, script)] + argv[2:]
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
