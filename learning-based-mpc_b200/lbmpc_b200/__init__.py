"""lbmpc_b200 — host-side mirror of the reference's (LB)MPC interface over the B200 engine.

The compute path is `liblbmpc_b200.so` (hand-written sm_100a CUDA behind the C ABI of
include/lbmpc.h); this package only packs arguments, mirrors the reference's function names
(`ocpLBMPC`, `ocpLMPC`, `mgcmDLTI`, `matOCP`, `getCONS`, `getCONSPOLY`) and shards batches
across ranks.  Nothing here computes a solve on the CPU.
"""
from .capi import (FORM, VARIANT, ST_OPTIMAL, ST_MAXITER, ST_INFEASIBLE, ST_NUMERICAL, LbmpcError, Solver,
                   load_library, pack_model, make_config, measure_fp64_peak)
from .model import (mgcmDLTI, matOCP, getCONS, getCONSPOLY, pdiff, tightened_state_set, moore_greitzer_model, double_integrator_model,
                    X_WP, U_WP)
from .drivers import ocpLBMPC, ocpLMPC, trueModel, transitionTrue, update_data
from . import sets, matio

__all__ = ["FORM", "VARIANT", "ST_OPTIMAL", "ST_MAXITER", "ST_INFEASIBLE", "ST_NUMERICAL", "LbmpcError", "Solver",
           "load_library", "pack_model", "make_config", "measure_fp64_peak", "mgcmDLTI", "matOCP", "getCONS", "getCONSPOLY",
           "pdiff", "tightened_state_set", "moore_greitzer_model", "double_integrator_model", "X_WP", "U_WP", "sets",
           "ocpLBMPC", "ocpLMPC", "trueModel", "transitionTrue", "update_data", "matio"]
