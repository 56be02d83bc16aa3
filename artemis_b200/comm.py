"""Cross-GPU ghost-zone exchange and the global dt reduction (one process per GPU).

Replaces, for blocks whose neighbour lives on another rank, the MPI path of
parthenon::SendBoundBufs / ReceiveBoundBufs / SetBounds (P:bvals/comms/
boundary_communication.cpp:48-334, P:utils/communication_buffer.hpp:209-420) and the
MPI_Allreduce(MIN) of EvolutionDriver::SetGlobalTimeStep (P:driver/driver.cpp:237).

Design.  MeshBlocks are partitioned block-spatially: rank r owns one
contiguous tile of the block lattice.  Where the reference sends one message per (block,
neighbour, variable) -- ~5k messages per stage at 256^3/GPU -- this sends ONE aggregated
buffer per peer per direction: the exchange runs as three sweeps (x1, x2, x3); sweep d sends
to each of the two face-neighbour ranks the `ng` innermost interior layers of every block on
that rank face, spanning the ENTIRE index range (ghosts included) of the other two
directions.  Edge and corner ghost zones therefore arrive through the face neighbours (the
x2 sweep forwards what the x1 sweep delivered), so a rank talks to at most 6 peers instead of
26 and no diagonal message exists.  Packing/unpacking is done by the descriptor-driven
kernels of the C ABI (ab200_halo_pack / ab200_halo_unpack, K8/K9) straight into / out of the
torch tensors handed to NCCL, whose send/recv go over NVLink (NVSwitch gives every peer full
bandwidth, so the three sweeps cost latency, not bandwidth: ~31 MB per stage per GPU).

The transport is torch.distributed (NCCL on GPUs; gloo in the CPU tests, where `backend`
is a numpy stand-in for the two kernels).  The planning code -- which blocks, which index
ranges, which peers -- is pure Python and is what the gloo tests cover.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .enums import Fluid

BC_NONE = 3  # AB200_BC_NONE: the face belongs to another rank


def rank_coords(rank, lay):
    return (rank % lay[0], (rank // lay[0]) % lay[1], rank // (lay[0] * lay[1]))


def rank_of(rc, lay):
    return rc[0] + lay[0] * (rc[1] + lay[1] * rc[2])


def ghost_var_runs(fluid_type, S):
    """Contiguous runs (var0, ncomp) of the FillGhost pack entries: gas prim rho, v, sie
    (pressure is not exchanged, src/gas/gas.cpp:243-270); dust prim rho, v."""
    if int(fluid_type) == int(Fluid.gas):
        return [(0, 4 * S), (5 * S, S)]
    return [(0, 4 * S)]


class SweepPlan:
    """Descriptors of one (direction, side) message: what this rank packs for the peer on
    that side and where the matching message from that peer is unpacked."""

    def __init__(self, peer, send, recv, nelem):
        self.peer = peer      # rank id
        self.send = send      # list of (fluid, block, var0, ncomp, si, ei, sj, ej, sk, ek, offset)
        self.recv = recv
        self.nelem = nelem    # doubles in the message


def plan_sweeps(mesh, fluids, lay, rl, periodic=(False, False, False)):
    """Per direction d: [lo-side SweepPlan or None, hi-side SweepPlan or None].

    mesh      this rank's UniformMesh (lattice_n = blocks of the tile)
    fluids    [(fluid_type, nspecies)] of the bound fluids
    lay, rl   rank lattice and this rank's coordinates in it
    periodic  whether the GLOBAL mesh is periodic in each direction (then the outermost
              ranks are each other's neighbours).
    """
    nbx, nby, nbz = mesh.lattice_n
    nbd = (nbx, nby, nbz)
    ng = mesh.ngd
    s = (mesh.is_, mesh.js, mesh.ks)
    e = (mesh.ie, mesh.je, mesh.ke)
    nt = (mesh.ni, mesh.nj, mesh.nk)
    plans = []
    for d in range(3):
        pair = [None, None]
        if nt[d] == 1 or lay[d] == 1 and not periodic[d]:
            plans.append(pair)
            continue
        for side in (0, 1):
            prc = list(rl)
            prc[d] += 1 if side else -1
            if prc[d] < 0 or prc[d] >= lay[d]:
                if not periodic[d]:
                    continue
                prc[d] %= lay[d]
            if lay[d] == 1:
                continue  # periodic wrap onto itself is a same-GPU exchange
            peer = rank_of(prc, lay)
            send, recv, off = [], [], 0
            # blocks on this rank face, lexicographic over the two tangential directions
            t1, t2 = [a for a in range(3) if a != d]
            for l2 in range(nbd[t2]):
                for l1 in range(nbd[t1]):
                    l = [0, 0, 0]
                    l[d] = nbd[d] - 1 if side else 0
                    l[t1], l[t2] = l1, l2
                    b = l[0] + nbx * (l[1] + nby * l[2])
                    lo_s = [0, 0, 0]
                    hi_s = [nt[0] - 1, nt[1] - 1, nt[2] - 1]
                    lo_r, hi_r = list(lo_s), list(hi_s)
                    if side:   # send the last ng interior layers, receive the upper ghosts
                        lo_s[d], hi_s[d] = e[d] - ng[d] + 1, e[d]
                        lo_r[d], hi_r[d] = e[d] + 1, e[d] + ng[d]
                    else:      # send the first ng interior layers, receive the lower ghosts
                        lo_s[d], hi_s[d] = s[d], s[d] + ng[d] - 1
                        lo_r[d], hi_r[d] = s[d] - ng[d], s[d] - 1
                    ncell = 1
                    for a in range(3):
                        ncell *= hi_s[a] - lo_s[a] + 1
                    for fl, S in fluids:
                        for var0, ncomp in ghost_var_runs(fl, S):
                            send.append((int(fl), b, var0, ncomp, lo_s[0], hi_s[0], lo_s[1],
                                         hi_s[1], lo_s[2], hi_s[2], off))
                            recv.append((int(fl), b, var0, ncomp, lo_r[0], hi_r[0], lo_r[1],
                                         hi_r[1], lo_r[2], hi_r[2], off))
                            off += ncomp * ncell
            pair[side] = SweepPlan(peer, send, recv, off)
        plans.append(pair)
    return plans


class PeerPlan:
    """One aggregated message per peer rank (direct scheme): every (block, neighbour offset)
    sub-box this rank owes that peer, and where the peer's matching message is unpacked."""

    def __init__(self, peer):
        self.peer = peer
        self.send = []     # (fluid, block, var0, ncomp, si, ei, sj, ej, sk, ek, offset)
        self.recv = []
        self.nsend = 0     # doubles
        self.nrecv = 0


def plan_direct(mesh, fluids, lay, rl, periodic=(False, False, False)):
    """Single-round exchange: for every neighbour offset o in {-1,0,1}^3 whose non-zero
    directions ALL cross onto another rank, the blocks on that rank face / edge / corner send
    their ng innermost layers along the non-zero directions and their INTERIOR range along the
    others, straight to the rank that owns the ghost zones (<= 26 peers, 7 in a 2x2x2 lattice).
    Ghost cells that combine a remote direction with a same-rank neighbour or a physical
    boundary are not sent at all: ab200_finish_remote_ghosts resolves them on the receiver from
    the delivered cells.  Returns [PeerPlan] sorted by peer.

    Matching: both ends walk the offsets in the same canonical order (the sender in o, the
    receiver in -o) and the blocks of a face / edge in lexicographic order of the tangential
    lattice coordinates, so the aggregated buffers line up without tags."""
    nbd = tuple(mesh.lattice_n)
    ng = mesh.ngd
    s = (mesh.is_, mesh.js, mesh.ks)
    e = (mesh.ie, mesh.je, mesh.ke)
    nt = (mesh.ni, mesh.nj, mesh.nk)
    offsets = [(ox, oy, oz) for oz in (-1, 0, 1) for oy in (-1, 0, 1) for ox in (-1, 0, 1)
               if (ox, oy, oz) != (0, 0, 0)]

    def peer_of(o):
        """rank owning the neighbour tile at offset o, or None if some non-zero direction does
        not cross onto another rank"""
        prc = list(rl)
        for d in range(3):
            if o[d] == 0:
                continue
            if nt[d] == 1 or lay[d] == 1:
                return None
            prc[d] += o[d]
            if prc[d] < 0 or prc[d] >= lay[d]:
                if not periodic[d]:
                    return None
                prc[d] %= lay[d]
        return rank_of(prc, lay)

    def boxes(o, sending):
        """(block, lo[3], hi[3]) of every block on the face / edge / corner of offset o"""
        rng = []
        for d in range(3):
            if o[d] == 0:
                rng.append(range(nbd[d]))
            else:
                rng.append([nbd[d] - 1] if o[d] > 0 else [0])
        out = []
        for lz in rng[2]:
            for ly in rng[1]:
                for lx in rng[0]:
                    l = (lx, ly, lz)
                    b = lx + nbd[0] * (ly + nbd[1] * lz)
                    lo, hi = [0, 0, 0], [0, 0, 0]
                    for d in range(3):
                        if o[d] == 0:
                            lo[d], hi[d] = s[d], e[d]
                        elif sending:
                            lo[d], hi[d] = ((e[d] - ng[d] + 1, e[d]) if o[d] > 0
                                            else (s[d], s[d] + ng[d] - 1))
                        else:
                            lo[d], hi[d] = ((e[d] + 1, e[d] + ng[d]) if o[d] > 0
                                            else (s[d] - ng[d], s[d] - 1))
                    out.append((b, lo, hi))
        return out

    plans = {}

    def add(o, sending):
        # the receiver walks the offsets in the sender's order: its offset is -o of the sender
        peer = peer_of(o)
        if peer is None:
            return
        p = plans.setdefault(peer, PeerPlan(peer))
        for b, lo, hi in boxes(o, sending):
            ncell = 1
            for d in range(3):
                ncell *= hi[d] - lo[d] + 1
            for fl, S in fluids:
                for var0, ncomp in ghost_var_runs(fl, S):
                    if sending:
                        p.send.append((int(fl), b, var0, ncomp, lo[0], hi[0], lo[1], hi[1], lo[2],
                                       hi[2], p.nsend))
                        p.nsend += ncomp * ncell
                    else:
                        p.recv.append((int(fl), b, var0, ncomp, lo[0], hi[0], lo[1], hi[1], lo[2],
                                       hi[2], p.nrecv))
                        p.nrecv += ncomp * ncell

    for o in offsets:
        add(o, True)
    for o in offsets:
        add(tuple(-v for v in o), False)
    return [plans[k] for k in sorted(plans)]


class DeviceBackend:
    """Pack/unpack through the C ABI into torch CUDA tensors."""

    def __init__(self, md):
        import torch
        self.torch = torch
        self.md = md
        self.device = torch.device(f"cuda:{md.device}")
        self._desc_cache = {}

    def alloc(self, n):
        return self.torch.empty(max(n, 1), dtype=self.torch.float64, device=self.device)

    def _descs(self, items, tensor):
        # plans and message buffers are persistent: build each ctypes descriptor array once
        key = (id(items), tensor.data_ptr())
        arr = self._desc_cache.get(key)
        if arr is None:
            arr = (capi.BndDesc * len(items))()
            base = tensor.data_ptr()
            for q, it in enumerate(items):
                arr[q] = capi.BndDesc(*it[:10], base + 8 * it[10])
            self._desc_cache[key] = arr
        return arr

    def pack(self, items, tensor):
        arr = self._descs(items, tensor)
        self.md.call("ab200_halo_pack", arr, len(items))

    def unpack(self, items, tensor):
        arr = self._descs(items, tensor)
        self.md.call("ab200_halo_unpack", arr, len(items))

    def allreduce_min(self, dist, value):
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    def time_state_tensor(self):
        """Zero-copy torch view of the library's device time state (dt, new_dt, time, ncycle)
        so the dt all-reduce runs on the device scalar with no host round trip."""
        if getattr(self, "_ts", None) is None:
            ptr = int(self.md.L.ab200_dt_device(self.md.ctx))

            class _View:
                __cuda_array_interface__ = {"shape": (4,), "typestr": "<f8", "data": (ptr, False),
                                            "version": 2, "strides": None}

            self._ts = self.torch.as_tensor(_View(), device=self.device)
        return self._ts

    def allreduce_min_device(self, dist):
        """new_dt <- min over ranks, in place on the device (P:driver/driver.cpp:237)."""
        dist.all_reduce(self.time_state_tensor()[1:2], op=dist.ReduceOp.MIN)


class HaloComm:
    """Remote part of AddBoundaryExchangeTasks + the dt all-reduce for one rank."""

    def __init__(self, md, lay, rl, rank, world, backend=None, periodic=(False, False, False),
                 dist=None):
        if dist is None:
            import torch.distributed as dist
        self.dist = dist
        self.md = md
        self.lay, self.rl, self.rank, self.world = tuple(lay), tuple(rl), rank, world
        self.backend = backend if backend is not None else DeviceBackend(md)
        fluids = [(ff.fp.fluid_type, ff.fp.nspecies) for ff in md.fluids]
        self.plans = plan_sweeps(md.mesh, fluids, self.lay, self.rl, periodic)
        self.bufs = []
        for pair in self.plans:
            row = []
            for p in pair:
                row.append(None if p is None else
                           (self.backend.alloc(p.nelem), self.backend.alloc(p.nelem)))
            self.bufs.append(row)
        self.bytes_per_exchange = 8 * sum(p.nelem for pair in self.plans for p in pair if p)
        # direct (single-round) scheme: ONE send slab and ONE receive slab, a slice per peer,
        # so all packing is one kernel launch, all unpacking another
        self.direct = plan_direct(md.mesh, fluids, self.lay, self.rl, periodic)
        ns = sum(p.nsend for p in self.direct)
        nr = sum(p.nrecv for p in self.direct)
        self.dsend, self.drecv = self.backend.alloc(ns), self.backend.alloc(nr)
        self._dsend_items, self._drecv_items, self._dslices = [], [], []
        so = ro = 0
        for p in self.direct:
            self._dsend_items += [it[:10] + (it[10] + so,) for it in p.send]
            self._drecv_items += [it[:10] + (it[10] + ro,) for it in p.recv]
            self._dslices.append((p.peer, self.dsend[so:so + p.nsend], self.drecv[ro:ro + p.nrecv]))
            so += p.nsend
            ro += p.nrecv
        self.bytes_per_direct_exchange = 8 * ns

    def pack_sweep(self, d):
        """ab200_halo_pack of both sides of direction d into the send buffers."""
        for side, p in enumerate(self.plans[d]):
            if p is not None:
                self.backend.pack(p.send, self.bufs[d][side][0])

    def unpack_sweep(self, d):
        """ab200_halo_unpack of both sides of direction d from the receive buffers."""
        for side, p in enumerate(self.plans[d]):
            if p is not None:
                self.backend.unpack(p.recv, self.bufs[d][side][1])

    def transfer_sweep(self, d):
        """One grouped send/recv per peer (NCCL over NVLink on GPUs)."""
        dist = self.dist
        pair = self.plans[d]
        sides = [s for s in (0, 1) if pair[s] is not None]
        ops = [dist.P2POp(dist.isend, self.bufs[d][side][0], pair[side].peer) for side in sides]
        # two ranks + periodic: both sides face the same peer, whose lo-side message is
        # our hi-side ghost data.  P2P ops to one peer match in posting order (NCCL has
        # no tags), so post the receives in the peer's send order.
        same_peer = len(sides) == 2 and pair[0].peer == pair[1].peer
        for side in (reversed(sides) if same_peer else sides):
            ops.append(dist.P2POp(dist.irecv, self.bufs[d][side][1], pair[side].peer))
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def exchange(self, md=None):
        """Three sweeps; call after the same-GPU exchange and before the physical BCs."""
        for d, pair in enumerate(self.plans):
            if pair[0] is None and pair[1] is None:
                continue
            self.pack_sweep(d)
            self.transfer_sweep(d)
            self.unpack_sweep(d)

    def transfer_direct(self):
        """One grouped send + recv per peer, all peers in one NCCL group."""
        dist = self.dist
        ops = [dist.P2POp(dist.isend, sb, peer) for peer, sb, _ in self._dslices]
        ops += [dist.P2POp(dist.irecv, rb, peer) for peer, _, rb in self._dslices]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def exchange_direct(self, md=None):
        """Single-round remote exchange (one pack launch, one NCCL group, one unpack launch).
        Delivers only the ghost cells whose non-interior directions are all remote; call
        ab200_finish_remote_ghosts afterwards (the device-resident cycle does)."""
        if not self._dslices:
            return
        self.backend.pack(self._dsend_items, self.dsend)
        self.transfer_direct()
        self.backend.unpack(self._drecv_items, self.drecv)

    # ---- overlapped form: the remote round runs on its own stream next to the local fill ----
    def _async_setup(self):
        import os
        t = getattr(self.backend, "torch", None)
        if (t is None or not t.cuda.is_available() or not isinstance(self.backend, DeviceBackend)
                or os.environ.get("AB200_NO_OVERLAP")):
            self._async = False
            return
        self._async = True
        self._t = t
        self.stream = t.cuda.Stream(device=self.backend.device)
        self._ev_stage = t.cuda.Event()
        self._ev_comm = t.cuda.Event()

    def begin_direct(self, md=None):
        """Start the single-round remote exchange of the stage that was just launched on the
        current stream.  On a GPU it runs on a second stream (pack, NCCL group, unpack), so
        ab200_fill_ghosts_local -- which reads interior zones only and writes none of the
        ghost cells the unpack writes -- can run concurrently; end_direct() joins."""
        if not hasattr(self, "_async"):
            self._async_setup()
        if not self._async or not self._dslices:
            self.exchange_direct(md)
            return
        t = self._t
        # the events below order the side stream against torch's CURRENT stream; the stage
        # kernels run on the stream given to ab200_create.  The two must be the same stream
        # (the C transport, NativeComm / ab200_comm_exchange_begin, orders against the context's
        # own stream and has no such requirement).
        cur = t.cuda.current_stream().cuda_stream
        if cur != getattr(self.md, "stream", 0):
            raise RuntimeError("HaloComm.begin_direct: torch's current stream is not the stream "
                               "the ab200 context was created on; use NativeComm or run under "
                               "torch.cuda.stream(ExternalStream(ctx_stream))")
        self._ev_stage.record(t.cuda.current_stream())
        self.md.call("ab200_set_halo_stream", C.c_void_p(self.stream.cuda_stream))
        try:
            with t.cuda.stream(self.stream):
                self.stream.wait_event(self._ev_stage)
                self.exchange_direct(md)
                self._ev_comm.record(self.stream)
        finally:
            self.md.call("ab200_set_halo_stream", None)

    def end_direct(self, md=None):
        if getattr(self, "_async", False) and self._dslices:
            self._t.cuda.current_stream().wait_event(self._ev_comm)

    def allreduce_min(self, value: float) -> float:
        """MPI_Allreduce(&dt, 1, MPI_DOUBLE, MPI_MIN) of P:driver/driver.cpp:237."""
        return self.backend.allreduce_min(self.dist, value)

    def allreduce_min_device(self):
        self.backend.allreduce_min_device(self.dist)


# ---------------------------------------------------------------------------------------------
# C-ABI transport (csrc/comm.cu): the planner, the NCCL communicator, the comm stream and the
# dt all-reduce live in libartemis_b200; Python only bootstraps the communicator (the 128-byte
# NCCL unique id travels over whatever the host already has -- torch.distributed here,
# MPI_Bcast in Parthenon).
# ---------------------------------------------------------------------------------------------
def c_plan_direct(L, mesh, fluids, lay, rl, periodic=(False, False, False)):
    """ab200_comm_plan_direct as [PeerPlan] (same structure as plan_direct, for the tests)."""
    I3 = C.c_int * 3
    nfl = len(fluids)
    rows = C.POINTER(C.c_longlong)()
    n = C.c_int(0)
    capi.check(L, L.ab200_comm_plan_direct(
        I3(*mesh.lattice_n), I3(mesh.ni, mesh.nj, mesh.nk), I3(mesh.is_, mesh.js, mesh.ks),
        I3(mesh.ie, mesh.je, mesh.ke), I3(*mesh.ngd), nfl,
        (C.c_int * max(nfl, 1))(*[int(f) for f, _ in fluids]),
        (C.c_int * max(nfl, 1))(*[int(s) for _, s in fluids]), I3(*lay), I3(*rl),
        I3(*[int(bool(p)) for p in periodic]), C.byref(rows), C.byref(n)), "ab200_comm_plan_direct")
    plans = {}
    try:
        for q in range(n.value):
            r = [int(rows[13 * q + c]) for c in range(13)]
            p = plans.setdefault(r[0], PeerPlan(r[0]))
            item = tuple(r[2:12]) + (r[12],)
            ncell = (r[7] - r[6] + 1) * (r[9] - r[8] + 1) * (r[11] - r[10] + 1)
            if r[1]:
                p.recv.append(item)
                p.nrecv = r[12] + r[5] * ncell
            else:
                p.send.append(item)
                p.nsend = r[12] + r[5] * ncell
    finally:
        L.ab200_comm_plan_free(rows)
    return [plans[k] for k in sorted(plans)]


class NativeComm:
    """One rank's handle on the library-owned transport.  Usage (see bench.py):

        comm = NativeComm(md, lay, rank, world, periodic)   # after bind + set_topology
        md.call("ab200_run_cycles_mr", integ, ncycles, tlim)
    """

    def __init__(self, md, lay, rank, world, periodic=(False, False, False), dist=None):
        import torch
        if dist is None:
            import torch.distributed as dist
        self.md = md
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            capi.check(md.L, md.L.ab200_comm_unique_id(idbuf), "ab200_comm_unique_id")
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).clone()
        dev = torch.device(f"cuda:{md.device}") if dist.get_backend() == "nccl" else torch.device("cpu")
        t = t.to(dev)
        dist.broadcast(t, src=0)
        idbytes = bytes(t.cpu().numpy().tobytes())
        md.call("ab200_comm_init", int(world), int(rank), C.c_char_p(idbytes))
        per = (C.c_int * 3)(*[int(bool(p)) for p in periodic])
        md.call("ab200_comm_set_layout", int(lay[0]), int(lay[1]), int(lay[2]), per)
        self.bytes_per_exchange = int(md.L.ab200_comm_bytes_per_exchange(md.ctx))

    def begin(self):
        self.md.call("ab200_comm_exchange_begin")

    def end(self):
        self.md.call("ab200_comm_exchange_end")

    def allreduce_min_device(self):
        ptr = int(self.md.L.ab200_dt_device(self.md.ctx))
        self.md.call("ab200_allreduce_min", C.c_void_p(ptr + 8))
