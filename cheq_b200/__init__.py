"""cheq_b200: B200-native QEq engine behind cheq's API (/root/reference/src/lib.rs:98-104).

The numerical path is CUDA only (cheq_b200/csrc, C ABI in include/cheq_b200.h); importing this package does
not need a GPU, creating a solver context does."""
from .errors import (CheqError, CudaError, DeserializationError, IoError, LinalgError, NoAtoms, NotConverged,
                     ParameterNotFound)
from .options import BasisType, DampingStrategy, SolverOptions
from .params import ElementData, Parameters, get_default_parameters
from .solver import Context, QEqSolver, default_context
from .types import Atom, AtomView, CalculationResult, ExternalPotential, PointCharge

__all__ = ["Atom", "AtomView", "BasisType", "CalculationResult", "CheqError", "Context", "CudaError",
           "DampingStrategy", "DeserializationError", "ElementData", "ExternalPotential", "IoError",
           "LinalgError", "NoAtoms", "NotConverged", "ParameterNotFound", "Parameters", "PointCharge",
           "QEqSolver", "SolverOptions", "default_context", "get_default_parameters"]
