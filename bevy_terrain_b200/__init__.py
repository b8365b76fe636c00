"""bevy_terrain_b200 — B200-native (sm_100a CUDA) implementation of bevy_terrain's hot path:
tile selection, vertex-grid generation and the tile-pyramid preprocessor, behind a C ABI
(include/bevy_terrain_b200.h).  This package is the thin host-side mirror of the reference's
plugin API; all compute lives in csrc/libbevy_terrain_b200.so and there is no CPU fallback.
"""
from . import _ffi
from ._ffi import (BT_FLAG_BLEND, BT_FLAG_HIGH_PRECISION, BT_FLAG_MORPH, BT_FLAG_NORMALS, BT_FORMAT_R16,
                   BT_FORMAT_RGBA8, BT_KIND_PLANAR, BT_KIND_SPHERICAL, BtError)


def __getattr__(name):  # lazy: importing the package must not require the built library
    import importlib
    if name.startswith("__"):
        raise AttributeError(name)
    api = importlib.import_module(__name__ + ".api")
    if name == "api":
        return api
    if hasattr(api, name):
        return getattr(api, name)
    raise AttributeError(name)
