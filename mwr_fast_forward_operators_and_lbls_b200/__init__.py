"""B200-native line-by-line microwave forward model: a drop-in for the PyRTlib `TbCloudRTE` path of
apschera2023uzk/MWR_fast_forward_operators_and_LBLs (python_src/proc/PyRTlib_processing.py).

    from mwr_fast_forward_operators_and_lbls_b200 import TbCloudRTE   # instead of pyrtlib.tb_spectrum
    from mwr_fast_forward_operators_and_lbls_b200 import simulate     # batched: (P,L) -> (P,F,A)

Host code is Python over PyTorch (device memory, streams, NCCL); every kernel is hand-written CUDA
for sm_100a in `csrc/`, reached through the C ABI of `include/mwrt.h` (libmwrt.so).  No Triton, no
multi-backend dispatch, no CPU fallback.  See DESIGN.md.
"""
from .models import implemented_models, get_model, load_descriptor, Model  # noqa: F401
from .forward import (Plan, simulate, vapor, absorption, slant_path, rte, bandpass,  # noqa: F401
                      measure_dfma_peak, pinned_fields, FIELDS)
from .tb_spectrum import TbCloudRTE  # noqa: F401
from .synth import synth_profiles, band_frequencies, dense_spectrum, HATPRO_FRQ, HATPRO_ELEV  # noqa: F401

__version__ = "0.1.0"
