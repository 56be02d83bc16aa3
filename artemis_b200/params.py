"""Input-deck handling: the subset of Parthenon's ParameterInput the hot path reads.

Mirrors the reference's deck semantics (``<block>`` headers, ``key = value  # comment``,
``block/key=value`` command-line overrides; P:parameter_input.cpp) and the selection logic
of Gas::Initialize / Dust::Initialize (src/gas/gas.cpp:40-208, src/dust/dust.cpp:34-100):
``gas/reconstruct``, ``gas/riemann``, ``gas/cfl``, ``gas/gamma``, floors, ``de_switch``,
``nspecies``, ``artemis/coordinates``, ``parthenon/time/integrator``, ``parthenon/mesh/nghost``.
"""
from __future__ import annotations

from dataclasses import dataclass

from .enums import Coordinates, CoordSelect, Fluid, ReconstructionMethod, RSolver


class ParameterInput:
    def __init__(self, text: str = "", overrides=()):
        self.blocks: dict[str, dict[str, str]] = {}
        block = None
        pending = ""
        for raw in text.splitlines():
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            if line.endswith("&"):
                pending += line[:-1].strip()
                continue
            line = pending + line
            pending = ""
            if line.startswith("<") and line.endswith(">"):
                block = line[1:-1].strip()
                self.blocks.setdefault(block, {})
                continue
            if "=" in line and block is not None:
                k, v = line.split("=", 1)
                self.blocks[block][k.strip()] = v.strip()
        for ov in overrides:
            path, v = ov.split("=", 1)
            blk, key = path.rsplit("/", 1)
            self.blocks.setdefault(blk, {})[key] = v.strip()

    @classmethod
    def from_file(cls, path, overrides=()):
        with open(path) as fh:
            return cls(fh.read(), overrides)

    def _get(self, block, key, default, conv):
        blk = self.blocks.setdefault(block, {})
        if key not in blk:
            if default is None:
                raise KeyError(f"Parameter <{block}>/{key} not found")
            blk[key] = str(default)
            return default
        return conv(blk[key])

    def GetOrAddReal(self, block, key, default):
        return self._get(block, key, default, float)

    def GetOrAddInteger(self, block, key, default):
        return self._get(block, key, default, lambda s: int(float(s)))

    def GetOrAddString(self, block, key, default):
        return self._get(block, key, default, str)

    def GetOrAddBoolean(self, block, key, default):
        return self._get(block, key, default, lambda s: s.lower() in ("true", "1"))

    def GetReal(self, block, key):
        return self._get(block, key, None, float)

    def GetInteger(self, block, key):
        return self._get(block, key, None, lambda s: int(float(s)))

    def GetString(self, block, key):
        return self._get(block, key, None, str)

    def DoesParameterExist(self, block, key):
        return key in self.blocks.get(block, {})


@dataclass
class FluidParams:
    """The StateDescriptor::Params the hot path reads (src/gas/gas.cpp:50-208)."""
    fluid_type: Fluid
    coords: Coordinates
    recon: ReconstructionMethod
    rsolver: RSolver
    cfl: float
    nspecies: int
    dfloor: float
    gamma: float = 5.0 / 3.0
    siefloor: float = 1.0e-20
    de_switch: float = 0.0

    @property
    def gm1(self) -> float:
        return self.gamma - 1.0

    @property
    def nvar(self) -> int:
        return (6 if self.fluid_type == Fluid.gas else 4) * self.nspecies


def _select_recon(name: str, nghost: int) -> ReconstructionMethod:
    need = {"pcm": 1, "plm": 2, "ppm": 3}
    if name not in need:
        raise ValueError("Reconstruction method not recognized.")
    if nghost < need[name]:
        raise ValueError(f"{name.upper()} requires at least {need[name]} ghost cells.")
    return ReconstructionMethod[name]


def gas_params(pin: ParameterInput, ndim: int) -> FluidParams:
    """Gas::Initialize selection logic, src/gas/gas.cpp:52-186."""
    nghost = pin.GetOrAddInteger("parthenon/mesh", "nghost", 2)
    coords = CoordSelect(pin.GetOrAddString("artemis", "coordinates", "cartesian"), ndim)
    recon = _select_recon(pin.GetOrAddString("gas", "reconstruct", "plm"), nghost)
    riemann = pin.GetOrAddString("gas", "riemann", "hllc")
    if riemann not in ("hllc", "hlle", "llf"):
        raise ValueError("Riemann solver (gas) not recognized.")
    eos = pin.GetOrAddString("gas", "eos", "ideal")
    if eos != "ideal":
        raise ValueError("only the ideal-gas EOS is on the hot path")
    return FluidParams(
        fluid_type=Fluid.gas, coords=coords, recon=recon, rsolver=RSolver[riemann],
        cfl=pin.GetOrAddReal("gas", "cfl", 0.8),
        nspecies=pin.GetOrAddInteger("gas", "nspecies", 1),
        dfloor=pin.GetOrAddReal("gas", "dfloor", 1.0e-20),
        gamma=pin.GetOrAddReal("gas", "gamma", 1.66666666667),
        siefloor=pin.GetOrAddReal("gas", "siefloor", 1.0e-20),
        de_switch=pin.GetOrAddReal("gas", "de_switch", 0.0))


def dust_params(pin: ParameterInput, ndim: int) -> FluidParams:
    """Dust::Initialize selection logic, src/dust/dust.cpp:44-100 (HLLC is gas-only)."""
    nghost = pin.GetOrAddInteger("parthenon/mesh", "nghost", 2)
    coords = CoordSelect(pin.GetOrAddString("artemis", "coordinates", "cartesian"), ndim)
    recon = _select_recon(pin.GetOrAddString("dust", "reconstruct", "plm"), nghost)
    riemann = pin.GetOrAddString("dust", "riemann", "hlle")
    if riemann not in ("hlle", "llf"):
        raise ValueError("Riemann solver (dust) not recognized.")
    return FluidParams(
        fluid_type=Fluid.dust, coords=coords, recon=recon, rsolver=RSolver[riemann],
        cfl=pin.GetOrAddReal("dust", "cfl", 0.8),
        nspecies=pin.GetOrAddInteger("dust", "nspecies", 1),
        dfloor=pin.GetOrAddReal("dust", "dfloor", 1.0e-20))
