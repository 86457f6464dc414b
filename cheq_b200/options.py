"""SolverOptions, BasisType, DampingStrategy (/root/reference/src/solver/options.rs:9-101)."""
from __future__ import annotations

import enum
from dataclasses import dataclass, field


class BasisType(enum.Enum):
    """options.rs:9-22; Sto is the default."""
    Gto = 0
    Sto = 1


@dataclass(frozen=True)
class DampingStrategy:
    """options.rs:26-45: None | Fixed(d) | Auto { initial }."""
    kind: str
    value: float = 0.0

    @staticmethod
    def none() -> "DampingStrategy":
        return DampingStrategy("none", 1.0)

    @staticmethod
    def fixed(d: float) -> "DampingStrategy":
        return DampingStrategy("fixed", float(d))

    @staticmethod
    def auto(initial: float = 0.4) -> "DampingStrategy":
        return DampingStrategy("auto", float(initial))

    # Rust-style aliases
    None_ = none
    Fixed = fixed
    Auto = auto


@dataclass(frozen=True)
class SolverOptions:
    """options.rs:53-100 (defaults at :90-100)."""
    tolerance: float = 1.0e-6
    max_iterations: int = 2000
    lambda_scale: float = 0.5
    hydrogen_scf: bool = True
    basis_type: BasisType = BasisType.Sto
    damping: DampingStrategy = field(default_factory=DampingStrategy.auto)

    @staticmethod
    def default() -> "SolverOptions":
        return SolverOptions()
