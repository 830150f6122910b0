"""Importable handle on the `mimamo-net_b200/` package (the directory name carries a hyphen, as
the project layout prescribes, so it cannot be imported by name).

    import mimamo_b200
    mimamo_b200.install()                       # puts mimamo-net_b200/api on sys.path
    from phase_difference_extractor import Phase_Difference_Extractor   # reference-style flat import
"""
import os
import sys

PACKAGE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mimamo-net_b200")
API_DIR = os.path.join(PACKAGE_DIR, "api")
CSRC_DIR = os.path.join(PACKAGE_DIR, "csrc")
LIB_PATH = os.path.join(PACKAGE_DIR, "libmimamo_b200.so")


def install():
    """Make the drop-in api/ modules importable by their reference names."""
    if API_DIR not in sys.path:
        sys.path.insert(0, API_DIR)
    return API_DIR
