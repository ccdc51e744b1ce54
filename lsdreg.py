"""Import alias: ``import lsdreg`` == the package directory ``lidar-slam-detection_b200/``
(a hyphen cannot appear in an ``import`` statement)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("lidar-slam-detection_b200")
