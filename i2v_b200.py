"""Importable name of the package directory `image-to-video-i2v-attack_b200/`.

The directory name is fixed by the build layout and is not a legal Python identifier, so this module
gives it one: it declares itself a package whose `__path__` is that directory, i.e.
`import i2v_b200.capi` loads `image-to-video-i2v-attack_b200/capi.py`.
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "image-to-video-i2v-attack_b200")]
__version__ = "0.1.0"
