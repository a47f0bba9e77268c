"""Importable alias for the hyphen-named package directory.

The product code lives in
`beyond-binary-fake-user-detection-a-credibility-aware-graph-based-recommender-system_b200/`
(a name Python cannot import directly); this shim points `credgcn.__path__` at it so
`import credgcn`, `from credgcn import graph, model, ...` work from the repo root.
"""
import pathlib as _pathlib

PACKAGE_DIR = (_pathlib.Path(__file__).resolve().parent.parent /
               "beyond-binary-fake-user-detection-a-credibility-aware-graph-based-recommender-system_b200")
__path__ = [str(PACKAGE_DIR)]
exec(compile((PACKAGE_DIR / "__init__.py").read_text(), str(PACKAGE_DIR / "__init__.py"), "exec"))
