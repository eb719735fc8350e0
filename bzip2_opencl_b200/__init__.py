"""bzip2_opencl_b200 -- B200-native (CUDA, sm_100a) bzip2 block compressor, drop-in for the compress path of
Stan1slav337/Bzip2-OpenCL.  The product is libbz2b200.so (csrc/, C ABI in include/bz2b200.h); `api` is its ctypes
binding and `include/` holds the reference-named C++ facade (OutputStream.hpp ...)."""
from . import corpus  # noqa: F401
from .api import Compressor, Bz2Error, load_library, stream_from_blocks  # noqa: F401
