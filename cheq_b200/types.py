"""Atoms, results and the external potential (/root/reference/src/types.rs:14-97, 136-292)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Protocol, Sequence, runtime_checkable


@runtime_checkable
class AtomView(Protocol):
    """types.rs:14-27: anything exposing an atomic number and a position (Angstrom).

    Python objects may expose them as attributes or as zero-argument methods."""
    atomic_number: int
    position: Sequence[float]


@dataclass(frozen=True)
class Atom:
    """types.rs:35-40."""
    atomic_number: int
    position: Sequence[float]


def atom_fields(atom):
    z = atom.atomic_number() if callable(getattr(atom, "atomic_number")) else atom.atomic_number
    p = atom.position() if callable(getattr(atom, "position")) else atom.position
    return int(z), (float(p[0]), float(p[1]), float(p[2]))


@dataclass
class CalculationResult:
    """types.rs:60-76."""
    charges: list
    equilibrated_potential: float
    iterations: int


@dataclass(frozen=True)
class PointCharge:
    """types.rs:84-97."""
    atomic_number: int
    position: Sequence[float]
    charge: float

    @staticmethod
    def new(atomic_number: int, position, charge: float) -> "PointCharge":
        return PointCharge(atomic_number, position, charge)


@dataclass
class ExternalPotential:
    """types.rs:136-292: point charges plus a uniform field (V/Angstrom); builder methods return self-copies."""
    _point_charges: list = field(default_factory=list)
    _uniform_field: tuple = (0.0, 0.0, 0.0)

    @staticmethod
    def new() -> "ExternalPotential":
        return ExternalPotential()

    @staticmethod
    def from_point_charges(charges) -> "ExternalPotential":
        return ExternalPotential(list(charges), (0.0, 0.0, 0.0))

    @staticmethod
    def from_uniform_field(field_) -> "ExternalPotential":
        return ExternalPotential([], tuple(float(c) for c in field_))

    def with_point_charges(self, charges) -> "ExternalPotential":
        return ExternalPotential(list(charges), self._uniform_field)

    def with_uniform_field(self, field_) -> "ExternalPotential":
        return ExternalPotential(list(self._point_charges), tuple(float(c) for c in field_))

    def point_charges(self):
        return self._point_charges

    def uniform_field(self):
        return list(self._uniform_field)

    def is_empty(self) -> bool:
        # types.rs:281-286
        return not self._point_charges and all(c == 0.0 for c in self._uniform_field)

    def num_point_charges(self) -> int:
        return len(self._point_charges)
