"""pathtracer-ocl_b200: B200-native replacement for the render path of eriklupander/pathtracer-ocl."""
__version__ = "0.1.0"
