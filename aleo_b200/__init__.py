"""aleo_b200 -- B200-native (sm_100a) BLS12-377 G1 MSM and Fr NTT behind the call signatures of
snarkVM 0.14.5's ``VariableBase::msm`` and ``EvaluationDomain`` (the proving hot path the Aleo SDK
reaches).  The compute lives in ``libaleo_b200.so`` (hand-written CUDA, C ABI in
include/aleo_b200.h); this package is the Python host-side mirror used by tests and bench.py.
There is no CPU fallback: without the built library and an sm_100 device every call raises."""
from ._lib import AleoB200Error, get_lib, NTT_FORWARD, NTT_INVERSE, NTT_STANDARD, NTT_COSET  # noqa: F401
from .domain import EvaluationDomain  # noqa: F401
from .msm import (VariableBase, gen_bases_dev, gen_scalars_dev, dlog_dot_dev, check_on_curve_dev,  # noqa: F401
                  AFFINE_STRIDE_RUST, AFFINE_STRIDE_PACKED, PROJECTIVE_BYTES)

from .kzg import ResidentSRS, KZG10  # noqa: F401,E402
from .poly import Evaluations, PolyMultiplier, field_op_dev  # noqa: F401,E402

__version__ = "0.1.0"
