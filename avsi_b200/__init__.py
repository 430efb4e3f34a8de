"""Importable alias of the package directory ``audio-visual-speech-inpainting_b200/``.

The task fixes the package directory name, which is not a valid Python identifier; this shim
puts that directory on ``avsi_b200.__path__`` so that ``import avsi_b200.models`` etc. resolve
to the files there.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         'audio-visual-speech-inpainting_b200')
__path__.insert(0, _PKG_DIR)
PACKAGE_DIR = _PKG_DIR
with open(_os.path.join(_PKG_DIR, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, '__init__.py'), 'exec'))
