"""ctypes prototypes of the forward-path entry points (filled in as the kernels land)."""
import ctypes as C


def declare(lib):
    pass
