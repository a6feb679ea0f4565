"""Loads the product package.  Its directory is named `grl-image-restoration_b200` (not a valid
Python identifier), so it is registered in sys.modules as `grl_image_restoration_b200`."""
import importlib.util
import os
import sys

PKG_NAME = "grl_image_restoration_b200"
PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "grl-image-restoration_b200")


def load_package():
    if PKG_NAME in sys.modules:
        return sys.modules[PKG_NAME]
    spec = importlib.util.spec_from_file_location(
        PKG_NAME, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR]
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules[PKG_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
