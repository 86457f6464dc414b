"""Element parameters and the TOML loader (/root/reference/src/params.rs:21-59, 89-133, 182-348)."""
from __future__ import annotations

import tomllib
from dataclasses import dataclass, field
from pathlib import Path

from .errors import DeserializationError, IoError

_SYMBOLS = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br "
            "Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er "
            "Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md "
            "No Lr Rf Db Sg Bh Hs Mt Ds Rg Cn Nh Fl Mc Lv Ts Og").split()
_SYMBOL_TO_Z = {s: i + 1 for i, s in enumerate(_SYMBOLS)}


def element_symbol_to_atomic_number(symbol: str):
    """params.rs:226-348 (case-sensitive, like the reference's match)."""
    return _SYMBOL_TO_Z.get(symbol)


@dataclass(frozen=True)
class ElementData:
    """params.rs:21-45.  TOML keys: chi, j, radius, n."""
    electronegativity: float
    hardness: float
    radius: float
    principal_quantum_number: int


@dataclass
class Parameters:
    """params.rs:52-59: `elements: HashMap<u8, ElementData>`."""
    elements: dict = field(default_factory=dict)

    @staticmethod
    def new() -> "Parameters":
        return Parameters()

    @staticmethod
    def load_from_file(path) -> "Parameters":
        try:
            content = Path(path).read_text()
        except OSError as e:  # params.rs:90-93
            raise IoError(path, e) from e
        return Parameters.load_from_str(content)

    @staticmethod
    def load_from_str(toml_str: str) -> "Parameters":
        try:
            doc = tomllib.loads(toml_str)
        except tomllib.TOMLDecodeError as e:
            raise DeserializationError(str(e)) from e
        if "elements" not in doc or not isinstance(doc["elements"], dict):
            raise DeserializationError("missing field `elements`")
        elements = {}
        for key, value in doc["elements"].items():
            # params.rs:198-205: key.parse::<u8>() or symbol lookup
            z = None
            if key.isascii() and key.isdigit() or (key.startswith("+") and key[1:].isdigit()):
                if 0 <= int(key) <= 255:
                    z = int(key)
            if z is None:
                z = element_symbol_to_atomic_number(key)
            if z is None:
                raise DeserializationError(f"invalid element key: '{key}'")
            if not isinstance(value, dict):
                raise DeserializationError(f"invalid type for element '{key}': expected a table")
            for name in ("chi", "j", "radius", "n"):
                if name not in value:
                    raise DeserializationError(f"missing field `{name}`")
            n = value["n"]
            if isinstance(n, bool) or not isinstance(n, int) or not (0 <= n <= 255):
                raise DeserializationError(f"invalid value for `n` of element '{key}': expected u8")
            for name in ("chi", "j", "radius"):
                if isinstance(value[name], bool) or not isinstance(value[name], (int, float)):
                    raise DeserializationError(f"invalid type for `{name}` of element '{key}': expected a float")
            elements[z] = ElementData(float(value["chi"]), float(value["j"]), float(value["radius"]), int(n))
        return Parameters(elements)


_DEFAULT = None


def get_default_parameters() -> Parameters:
    """lib.rs:127-133: the embedded Rappe-Goddard table, parsed once and cached."""
    global _DEFAULT
    if _DEFAULT is None:
        path = Path(__file__).resolve().parent / "data" / "qeq_params.toml"
        _DEFAULT = Parameters.load_from_str(path.read_text())
    return _DEFAULT
