"""Enumerations shared by the host mirror and the C ABI (values match include/ab200.h).

Order follows the reference's ``enum class`` declarations, src/artemis.hpp:78-105.
"""
from enum import IntEnum


class Coordinates(IntEnum):
    cartesian = 0
    cylindrical = 1
    spherical1D = 2
    spherical2D = 3
    spherical3D = 4
    axisymmetric = 5


class RSolver(IntEnum):
    hllc = 0
    hlle = 1
    llf = 2


class ReconstructionMethod(IntEnum):
    pcm = 0
    plm = 1
    ppm = 2


class Fluid(IntEnum):
    gas = 0
    dust = 1


class BoundaryFlag(IntEnum):
    periodic = 0
    outflow = 1
    reflect = 2
    # user condition that depends on position only (`ic` of src/pgen/disk.hpp, strat.hpp):
    # AB200_BC_FIXED -- the ghost zones keep the problem generator's profile
    fixed = 4
    # state-dependent user conditions of the shearing-box problem generators
    # (src/pgen/strat.hpp:154-666, inputs/ssheet/ssheet.in): AB200_BC_EXTRAP on x1 / x3 faces,
    # AB200_BC_INFLOW on x2 faces; applied per block through ab200_block_bcs
    extrap = 5
    inflow = 6


def CoordSelect(sys: str, ndim: int) -> Coordinates:
    """geometry::CoordSelect, src/geometry/geometry.hpp:37-56."""
    if sys == "cartesian":
        return Coordinates.cartesian
    if sys == "spherical":
        return (Coordinates.spherical1D, Coordinates.spherical2D,
                Coordinates.spherical3D)[min(ndim, 3) - 1]
    if sys == "cylindrical":
        return Coordinates.cylindrical
    if sys == "axisymmetric":
        return Coordinates.axisymmetric
    raise ValueError(f"Coordinate type not recognized: {sys!r}")


# (gam0, gam1, beta) per stage: P:time_integration/low_storage_integrator.cpp:50-130
INTEGRATORS = {
    "rk1": ((0.0, 1.0, 1.0),),
    "rk2": ((0.0, 1.0, 1.0), (0.5, 0.5, 0.5)),
    "vl2": ((0.0, 1.0, 0.5), (0.0, 1.0, 1.0)),
    "rk3": ((0.0, 1.0, 1.0), (0.25, 0.75, 0.25), (2.0 / 3.0, 1.0 / 3.0, 2.0 / 3.0)),
}
