"""CPU oracle for the packed QuantLinear hot path of xvyaward/qeft.

TEST INFRASTRUCTURE ONLY.  Nothing under ``qeft_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs do, and only as the checker / the timed CPU baseline.

Parity status: the reference ships no tests, golden vectors or fixtures for this
path, so the oracle is pinned against outputs of the reference's
own Python (``pack_intweight``, ``pack_oweight``, ``QuantLinear.pack``,
``sparse_to_dense_ids``, ``lm_pack``/``save_model``) imported in the build
container; the vectors live in ``tests/golden/`` with ``make_golden.py``.
The reference has no CPU implementation of the forward arithmetic (every
``forward_*`` calls the CUDA extension); the forward oracle is pinned against the
reference's OWN kernels run on the B200 (``oracle/build_ref.py`` ->
``oracle/_ref/qeft_cuda_ref.so``, ``tests/test_reference_gpu.py``).  The backward
oracle restates the math BASELINE.json defines (the reference's backward is not
runnable) and is "parity unpinned" by reference-run outputs.
"""
from .qeft_oracle import *  # noqa: F401,F403
