"""Loads the package directory `cnn-super-resolution_b200/` (its name, fixed by the project
layout, is not a Python identifier) under the importable name `cnn_super_resolution_b200`."""
import importlib.util
import os
import sys

NAME = "cnn_super_resolution_b200"
PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cnn-super-resolution_b200")


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(
        NAME, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
