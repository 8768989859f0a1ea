"""qeft_b200 -- B200-native (sm_100a) packed QuantLinear for QEFT checkpoints.

Drop-in for the reference's ``qeft/qlinear.py`` module API and its ``qeft_cuda`` extension
(``gemm_4bit``, ``gemv_4bit``, ``gemv_4bit_qeft``).  All compute goes through the C-ABI library
``qeft_b200/csrc/libqeft_b200.so`` (include/qeft_b200.h); there is no CPU or eager fallback.
"""
from . import qeft_cuda  # noqa: F401
from .qlinear import (QuantLinear, QuantMatMul, QuantMatMulQEFT, pack_intweight, pack_oweight,  # noqa: F401
                      unpack_intweight)
from .reorder import sparse_to_dense_ids  # noqa: F401
from .quant import make_quant, lm_pack  # noqa: F401

__version__ = "0.1.0"
