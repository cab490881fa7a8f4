"""Import alias: ``import b200dt`` == the package directory
``yolo---small-target-recognition---kalman-trajectory-prediction_b200`` (whose name, fixed by the
project layout, is not a Python identifier)."""
import importlib
import sys

_pkg = importlib.import_module("yolo---small-target-recognition---kalman-trajectory-prediction_b200")
sys.modules[__name__] = _pkg
