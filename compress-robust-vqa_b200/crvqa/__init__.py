"""ctypes binding of libcrvqa.so (the sm_100a kernels) plus the torch glue that stands where
``_Binarizer1`` + ``F.linear`` / ``torch.kthvalue`` / the loss graphs stand in the reference.

Nothing in this package falls back to CPU or to stock PyTorch math for the hot ops: if the shared
library is missing or the device is not CUDA the calls raise.
"""
from ._lib import lib, check, CrvqaError, LIB_PATH  # noqa: F401
