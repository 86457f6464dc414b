"""CheqError variants (/root/reference/src/error.rs:18-68), same names and messages."""
from __future__ import annotations


class CheqError(Exception):
    """Base class; mirrors `enum CheqError`."""


class ParameterNotFound(CheqError):
    def __init__(self, atomic_number: int):
        self.atomic_number = int(atomic_number)
        super().__init__(f"Atomic parameters not found for element with atomic number: {self.atomic_number}")


class NotConverged(CheqError):
    def __init__(self, max_iterations: int, delta: float):
        self.max_iterations = int(max_iterations)
        self.delta = float(delta)
        super().__init__(f"SCF failed to converge after {self.max_iterations} iterations. "
                         f"Final charge difference: {self.delta:.2e}")


class LinalgError(CheqError):
    def __init__(self, message: str):
        self.message = message
        super().__init__(f"Failed to solve the linear matrix system: {message}")


class IoError(CheqError):
    def __init__(self, path, source: Exception):
        self.path = path
        self.source = source
        super().__init__(f"I/O error at path '{path}': {source}")


class DeserializationError(CheqError):
    def __init__(self, message: str):
        super().__init__(f"Failed to deserialize TOML parameters: {message}")


class NoAtoms(CheqError):
    def __init__(self):
        super().__init__("Input validation failed: at least one atom is required for a calculation")


class CudaError(CheqError):
    """Device/runtime failure; no reference equivalent (the reference has no device)."""
