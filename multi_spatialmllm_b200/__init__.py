"""B200-native MultiSPA geometry hot path.

Sub-modules:
  _lib     ctypes binding of libmsp.so (C ABI in include/msp.h) — raises if the library is not built
  engine   DeviceScene + per-scene workers (needs a CUDA device)
  synth    seeded synthetic scenes / videos (host only)
  handler  drop-in SceneInfoHandler / project_points surface of the reference
"""
__all__ = ["_lib", "engine", "synth", "handler"]
__version__ = "0.1.0"
