"""openshadinglanguage_b200 — B200-native execution back end for OSL's
data-parallel hot path (a compiled ShaderGroup over a large SoA batch of
shading points).

The product is the C-ABI shared library `libosl_b200.so` (see include/osl_b200.h);
this package is the thin Python binding used by the tests and bench.py, plus
the host-side mirror of testshade's grid setup.  It FAILS LOUDLY if the CUDA
library is missing: there is no CPU fallback in the product path.
"""
from .api import (B200Error, ShaderGroup, lib, library_path, shadeop_hash,  # noqa: F401
                  shadeop_noise, SG_FIELDS, launch_count, add_texture, pack_userdata)
from .testshade import grid_globals  # noqa: F401

__all__ = ["B200Error", "ShaderGroup", "lib", "library_path", "shadeop_noise", "shadeop_hash",
           "grid_globals", "SG_FIELDS", "launch_count", "add_texture", "pack_userdata"]
