"""mpstime.jl_b200 -- B200-native (sm_100a) implementation of the MPSTime.jl hot path
(fitMPS two-site sweep, classify, MPS_impute) behind the reference's own interface.

The numeric path is libmpstime_b200.so (hand-written CUDA, C ABI in include/mpstime_b200.h);
this package is the host-side mirror of the Julia API.  There is no CPU fallback."""
from .core import Context, MPSTError, make_opts, BASIS_IDS, TIMER_NAMES      # noqa: F401
from .api import (MPSOptions, TrainedMPS, EncodedTimeSeriesSet, fitMPS, classify,                # noqa: F401
                  ImputationProblem, init_imputation_problem, get_predictions_batch, MPS_impute,
                  MPSClassifier, make_grid)
from .preprocess import (transform_train_data, transform_test_data, invert_test_transform,       # noqa: F401
                         generate_starting_mps, sort_by_class)
from . import dist                                                                                # noqa: F401
from ._lib import LIB_PATH, SIGNATURES                                                            # noqa: F401
