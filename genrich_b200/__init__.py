"""genrich_b200: B200-native pileup -> p-value -> q-value -> peak engine behind the
Genrich interface.  The compute lives in csrc/libgenrich_cuda.so (hand-written
sm_100a CUDA behind the C-ABI of include/genrich_cuda.h); this package is the
thin host-side mirror used by the tests, the bench and the multi-GPU launcher."""
from .capi import (Api, Context, GenrichError, GrParams, load_cuda, make_params,  # noqa: F401
                   PEAK_DTYPE, ABI_SYMBOLS, CUDA_LIB)
from .host import (fragments_to_intervals, run_replicates, format_narrowpeak,  # noqa: F401
                   format_log, lpt_shard)
