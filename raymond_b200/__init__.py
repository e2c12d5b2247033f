"""raymond_b200 — B200-native (sm_100a) implementation of raymond's path-tracing hot path.

    raymond_b200.api       Python host mirror of the reference API over the C ABI (include/raymond.h)
    raymond_b200.fixtures  benchmark scenes and mesh generators (harness data)
    raymond_b200.build     builds libraymond_cuda.so in-tree with nvcc

Names of `api` are re-exported lazily so that importing the package (e.g. for the fixtures) does not
load the CUDA library.
"""
__all__ = ["api", "fixtures", "build"]


def __getattr__(name):
    if name in ("api", "fixtures", "build", "distributed"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    from . import api
    try:
        return getattr(api, name)
    except AttributeError:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}") from None
