"""Import alias for the package directory `drop-clip_b200/`.

The directory name follows the project naming contract and is not a valid Python
identifier, so this one-file shim turns itself into a package whose search path is that
directory: `import dropclip_b200.feature_fusion` loads `drop-clip_b200/feature_fusion.py`.
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "drop-clip_b200")]
__version__ = "0.1.0"
PACKAGE_DIR = __path__[0]
