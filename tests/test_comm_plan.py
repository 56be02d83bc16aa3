"""The single-round exchange planner of the C ABI (csrc/comm.cu: ab200_comm_plan_direct, host
code, no GPU needed) against the Python planner the gloo tests cover (comm.py:plan_direct):
same peers, same descriptors, same offsets, row for row, on every rank of several lattices."""
import itertools

import pytest

from artemis_b200 import capi
from artemis_b200.comm import c_plan_direct, plan_direct, rank_coords
from artemis_b200.enums import Fluid
from artemis_b200.mesh import UniformMesh


@pytest.mark.parametrize("lay,periodic", [
    ((2, 1, 1), (False, False, False)), ((2, 2, 1), (False, True, False)),
    ((2, 2, 2), (False, False, False)), ((2, 2, 2), (True, True, True)),
    ((4, 2, 1), (True, False, False)), ((1, 3, 2), (False, True, True)),
])
@pytest.mark.parametrize("ndim", [2, 3])
def test_c_planner_equals_python_planner(lay, periodic, ndim):
    L = capi.load("fast")
    if ndim == 2 and lay[2] > 1:
        pytest.skip("2-D mesh has one rank layer in x3")
    tile = (2, 3, 2 if ndim == 3 else 1)          # blocks per rank tile
    bnx = (8, 6, 4 if ndim == 3 else 1)
    nx = tuple(tile[d] * bnx[d] * lay[d] for d in range(3))
    fluids = [(Fluid.gas, 2), (Fluid.dust, 3)]
    world = lay[0] * lay[1] * lay[2]
    for rank in range(world):
        rl = rank_coords(rank, lay)
        mesh = UniformMesh(nx=nx, xmin=(0, 0, 0), xmax=(1, 1, 1), block_nx=bnx, nghost=4,
                           lattice_lo=tuple(rl[d] * tile[d] for d in range(3)), lattice_n=tile)
        want = plan_direct(mesh, fluids, lay, rl, periodic)
        got = c_plan_direct(L, mesh, fluids, lay, rl, periodic)
        assert [p.peer for p in got] == [p.peer for p in want]
        for g, w in zip(got, want):
            assert g.send == w.send and g.recv == w.recv
            assert (g.nsend, g.nrecv) == (w.nsend, w.nrecv)


def test_send_and_receive_sides_line_up_across_ranks():
    """What rank a sends to rank b is, message for message, what b expects from a."""
    L = capi.load("fast")
    lay, tile, bnx = (2, 2, 2), (2, 2, 2), (8, 8, 8)
    nx = tuple(tile[d] * bnx[d] * lay[d] for d in range(3))
    fluids = [(Fluid.gas, 1)]
    plans = {}
    for rank in range(8):
        rl = rank_coords(rank, lay)
        mesh = UniformMesh(nx=nx, xmin=(0, 0, 0), xmax=(1, 1, 1), block_nx=bnx, nghost=4,
                           lattice_lo=tuple(rl[d] * tile[d] for d in range(3)), lattice_n=tile)
        plans[rank] = {p.peer: p for p in c_plan_direct(L, mesh, fluids, lay, rl)}
    for a, b in itertools.permutations(range(8), 2):
        assert plans[a][b].nsend == plans[b][a].nrecv
        sizes_a = [(s[3], s[5] - s[4], s[7] - s[6], s[9] - s[8]) for s in plans[a][b].send]
        sizes_b = [(r[3], r[5] - r[4], r[7] - r[6], r[9] - r[8]) for r in plans[b][a].recv]
        assert sizes_a == sizes_b
