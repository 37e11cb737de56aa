"""univs_b200 -- B200 (sm_100a) native implementation of the UniVS per-clip forward hot path.

Only what the path needs lives here: `csrc/` (CUDA kernels + the C ABI of include/univs_b200.h),
`_cabi.py` (ctypes binding of that ABI), `ops.py` (tensor-level operator wrappers) and the host-side
mirror of the reference's module interface (`modeling/`, `meta_arch.py`).  There is no CPU fallback:
every operator raises if the CUDA library is missing or an input is not a CUDA tensor.
"""
__version__ = "0.1.0"
