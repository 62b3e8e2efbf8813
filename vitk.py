"""Importable alias for the hyphenated package directory:  `import vitk`."""
import importlib
import sys
from pathlib import Path

_root = str(Path(__file__).resolve().parent)
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("automated-recycling-sorter-with-vision-transformers_b200")
sys.modules[__name__] = _pkg
