"""b200cam - B200-native (sm_100a) differentiable optical-encoder camera.

Drop-in replacements for the reference's camera modules

* ``Camera``         <- ``Face-DeId/Camera/Optics.py:9``
* ``OpticsZernike``  <- ``Image_Caption/Camera/Lens.py:11``

whose hot path (height map -> pupil phase -> PSF by FFT propagation -> FFT convolution
with the image -> sensor normalisation, and the backward pass into the height map) runs
in hand-written CUDA kernels behind the C-ABI library ``libb200cam.so``
(``include/b200cam.h``).  There is no CPU fallback: every compute entry point raises if
the library or a CUDA device is missing.
"""
__version__ = "0.1.0"

from . import zernike, synthetic  # noqa: F401  (pure-python helpers, CPU-safe)


def __getattr__(name):
    # heavy / GPU-facing modules are imported lazily so that CPU-only tooling can import the package
    if name in ("Camera",):
        from .optics import Camera
        return Camera
    if name in ("OpticsZernike",):
        from .lens import OpticsZernike
        return OpticsZernike
    if name in ("lib", "load_library"):
        from . import _lib
        return getattr(_lib, name)
    raise AttributeError(name)
