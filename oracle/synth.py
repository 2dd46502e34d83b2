"""TEST INFRASTRUCTURE shim.  The deterministic synthetic weight / batch generator moved into the package
(`gw-depth_b200/synth.py`): `bench.py`'s product arm needs synthetic inputs and must not import anything under `oracle/`.  The
oracle, the golden-fixture script and the tests keep importing `synth` from here; the generated tensors are unchanged."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_impl = importlib.import_module("gw-depth_b200.synth")
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
