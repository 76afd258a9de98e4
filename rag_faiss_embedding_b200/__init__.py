"""Importable alias of the package directory `rag-faiss-embedding_b200/` (a hyphen is not a valid
Python identifier).  All code lives there; this module only redirects the import."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rag-faiss-embedding_b200")
__path__ = [_impl]
with open(_os.path.join(_impl, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, "__init__.py"), "exec"))
