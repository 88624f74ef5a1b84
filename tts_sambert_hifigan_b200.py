"""Import shim: the package directory is named `tts-sambert_hifigan_b200/` (with
a hyphen, as the build contract names it), which Python cannot import directly.
`import tts_sambert_hifigan_b200` loads that directory as a regular package under
the underscore name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tts-sambert_hifigan_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
