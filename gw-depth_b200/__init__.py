"""gw-depth_b200: B200-native (sm_100a) implementation of the GW-Depth model forward hot path.

The directory name follows the repo contract and is not a Python identifier; import it with
`importlib.import_module("gw-depth_b200")` or through the root-level alias module `gwdepth_b200`.
"""
__version__ = "0.1.0"
