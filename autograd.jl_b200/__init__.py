"""agb200 -- B200-native array hot path behind AutoGrad.jl's operator API (host mirror).

Surface (same names and meaning as the reference, src/AutoGrad.jl:2-5): Param, diff (`@diff`), grad,
gradloss, value, params, primitive / zerograd (`@primitive` / `@zerograd`), plus the device array
type DArray and the data-parallel helpers.  Everything numeric runs in libagb200.so (sm_100a);
importing this package without the built library raises ImportError -- there is no CPU fallback.
"""
from . import _lib
from ._lib import DeviceError, DimensionMismatch, UnsupportedOnDevice
from .darray import (DArray, IArray, device_info, gemm_engine, init, launch_count, set_gemm_engine, sync,
                     zeros)
from .core import (Node, Param, Result, Tape, Tracked, Value, differentiate, diff, gcnode, grad,
                   gradloss, params, recording, release_node, set_gc_function, set_timer, timer_report, value)
from .sparse import Sparse, flat_dest, to_indices
from .rules import *  # noqa: F401,F403  (the primitive catalogue)
from .rules import (Primitive, abs_, addto, back, full, max_, min_, pow_, primitive, sum_, zerograd)
from . import rules, workloads
from .graph import StepGraph, Event, pinned_empty
from .comm import DataParallel
from .gradcheck import gradcheck
from .vm import E, Expr, primitive_expr
from .lazy import set_fusion, set_col_fusion
from .ctape import CTape
from .loader import CharBatches, StagedInput, UByteBatch, charlm_minibatch_ids, read_chars, read_idx_ubyte

__all__ = [n for n in dir() if not n.startswith("_")]
