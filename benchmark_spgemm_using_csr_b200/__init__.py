"""bhsparse-b200: B200-native CSR SpGEMM (C = A*B) behind the bhSPARSE class API.

Host side only here; the compute path is the CUDA library built from csrc/
(`python -m benchmark_spgemm_using_csr_b200.build`).  Importing the package
never touches the GPU; using it without the library or without an sm_100
device raises -- there is no CPU fallback.
"""
from . import capi, generators  # noqa: F401
from .bhsparse import BHSPARSE_CUDA, BHSPARSE_SUCCESS, NUM_PLATFORMS, bhsparse, spgemm  # noqa: F401

__version__ = "0.1.0"
