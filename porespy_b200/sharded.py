"""Volumes sharded as z-slabs over the GPUs of one box (one process per GPU, torch.distributed).

Rank r owns the planes [z0_r, z1_r) of a global [nz][ny][nx] volume (C order: a slab is one
contiguous block).  What is slab-local and what is exchanged:

* EDT x and y passes: local.  z pass: slab -> pencil all-to-all (`all_to_all_single`; NCCL over
  NVLink on GPUs), z pass on the pencil [nz][ny/P][nx], all-to-all back.  The pack for the first
  transpose is fused into the y-pass store (`psb200_edt_xy_u8(ysplit)`), the pencil is the receive
  buffer itself, and the way back sends contiguous z ranges.
* max d2 (for `sizes=int`, F:1131-1132): one scalar all-reduce.
* per radius (F:1177-1209): everything is local except the reach of the ball across the slab
  faces, W = ceil(sqrt(T)) - 1 planes: the byte pipeline exchanges W planes of the reach map, the
  bit pipeline W planes of seed bits, with the two z-neighbours only (send/recv).
* access-limited flooding (F:1181-1183): local union-find plus an exchange of the reached flags of
  the slab faces, repeated until no rank reaches anything new.

The same driver runs on CPU tensors over gloo with a numpy backend (tests/cpu_backend.py) so the
partitioning and exchange logic is covered without a GPU; the product backend below only ever
calls libpsb200.so.
"""
import ctypes

import numpy as np

from . import _device as dev
from . import _host as host
from . import _lib


def _dist():
    import torch.distributed as dist
    return dist


def split_counts(n, parts):
    """Balanced contiguous partition of range(n) into `parts` pieces (first pieces one longer)."""
    q, r = divmod(int(n), int(parts))
    return [q + (1 if i < r else 0) for i in range(parts)]


def ceil_split_counts(n, parts):
    """Partition with a fixed stride ceil(n/parts) (the layout `psb200_edt_xy_u8(ysplit)` writes)."""
    s = -(-int(n) // int(parts))
    return s, [max(0, min(s, n - d * s)) for d in range(parts)]


class CudaBackend:
    """Step primitives on one GPU: thin calls into libpsb200.so (torch tensors = device buffers)."""

    def __init__(self, ctx):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device)
        self.bit_tmax = 200
        self.uf_records = True      # step-level flood on the row-rooted forest + link records (False: per-voxel job lists)

    # -- allocation
    def empty(self, n, dtype):
        return self.torch.empty(int(n), dtype=dtype, device=self.device)

    def zeros(self, n, dtype):
        return self.torch.zeros(int(n), dtype=dtype, device=self.device)

    def to_u8(self, arr, positive=False):
        return dev.to_device_u8(arr, self.ctx, positive=positive).reshape(-1)

    def bit_ok(self, shape, T):
        return shape[2] % 32 == 0 and T <= self.bit_tmax

    # -- EDT
    def edt_xy(self, im_u8, shape, ysplit):
        nz, ny, nx = shape
        h = self.empty(nz * ny * nx, self.torch.int32)
        ws = self.ctx.workspace(2 * nz * ny * nx + 1024)
        _lib.check(self.ctx.lib.psb200_edt_xy_u8(self.ctx.handle, dev.ptr(im_u8), dev.ptr(h), nz, ny, nx,
                                                 int(ysplit), dev.ptr(ws), ws.numel(), dev.stream_ptr()))
        return h

    def edt_z(self, h, shape):
        nz, ny, nx = shape
        out = self.empty(nz * ny * nx, self.torch.int32)
        mx = self.empty(1, self.torch.int32)
        if nz * ny * nx == 0:
            return out, 0
        _lib.check(self.ctx.lib.psb200_edt_z_u32(self.ctx.handle, dev.ptr(h), dev.ptr(out), 0, dev.ptr(mx),
                                                 nz, ny, nx, dev.stream_ptr()))
        return out, int(mx.cpu().numpy().view(np.uint32)[0])

    def edt_ext(self, ext_u8, shape_ext, lo, nzl):
        """Exact EDT of the extended slab (own planes plus input halo planes), one-GPU kernels; returns the own
        planes [lo, lo + nzl) as a flat int32 tensor and their maximum."""
        nze, ny, nx = shape_ext
        if nze == 1:
            d2e, mx = dev.edt_run(self.ctx, ext_u8, shape_ext, want_max=True)
        else:
            d2e, mx = dev.edt_run(self.ctx, ext_u8, shape_ext, want_max=True, zmax=(lo, lo + nzl))
        return d2e[lo * ny * nx:(lo + nzl) * ny * nx], mx

    # -- per-radius steps
    def classify(self, d2, T):
        cls = self.empty(d2.numel(), self.torch.uint8)
        T = np.ascontiguousarray(T, dtype=np.uint32)
        _lib.check(self.ctx.lib.psb200_lt_classify(self.ctx.handle, dev.ptr(d2),
                                                   T.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), len(T),
                                                   dev.ptr(cls), d2.numel(), dev.stream_ptr()))
        return cls

    def lt_xy(self, cls, k, T, shape):
        nz, ny, nx = shape
        reach = self.empty(nz * ny * nx, self.torch.uint8)
        n = nz * ny * nx
        ws = self.ctx.workspace(n + n // 8 + n // 32 + 8192)     # x-distance bytes + seed bits + activity flags of the radius
        _lib.check(self.ctx.lib.psb200_lt_xy(self.ctx.handle, dev.ptr(cls), int(k), int(T), dev.ptr(reach),
                                             nz, ny, nx, dev.ptr(ws), ws.numel(), dev.stream_ptr()))
        return reach

    def lt_z(self, reach, m_lo, m_hi, idx, k, T, shape):
        nz, ny, nx = shape
        nlo = 0 if m_lo is None else m_lo.numel() // (ny * nx)
        nhi = 0 if m_hi is None else m_hi.numel() // (ny * nx)
        _lib.check(self.ctx.lib.psb200_lt_z(self.ctx.handle, dev.ptr(reach), dev.ptr(m_lo), nlo, dev.ptr(m_hi),
                                            nhi, dev.ptr(idx), int(k), int(T), nz, ny, nx, dev.stream_ptr()))

    def mask_pack(self, src_u8):
        """0 / non-zero bytes -> bits (psb200_mask_pack_u8)."""
        n = src_u8.numel()
        bits = self.empty((n + 7) // 8, self.torch.uint8)
        _lib.check(self.ctx.lib.psb200_mask_pack_u8(self.ctx.handle, dev.ptr(src_u8), dev.ptr(bits), n, dev.stream_ptr()))
        return bits

    def mask_unpack(self, bits, dst_u8):
        """bits -> 0 / 1 bytes into dst_u8 (psb200_mask_unpack_u8)."""
        _lib.check(self.ctx.lib.psb200_mask_unpack_u8(self.ctx.handle, dev.ptr(bits), dev.ptr(dst_u8), dst_u8.numel(),
                                                      dev.stream_ptr()))

    def halo_cone(self, reach, shape, depth, side):
        """One plane for a z-neighbour: the cone value its sweep receives through the shared face (psb200_lt_halo_cone)."""
        nz, ny, nx = shape
        out = self.empty(ny * nx, self.torch.uint8)
        _lib.check(self.ctx.lib.psb200_lt_halo_cone(self.ctx.handle, dev.ptr(reach), nz, ny, nx, int(depth), int(side),
                                                    dev.ptr(out), dev.stream_ptr()))
        return out

    def pack(self, cls, k, out_bits, shape):
        nz, ny, nx = shape
        _lib.check(self.ctx.lib.psb200_lt_pack(self.ctx.handle, dev.ptr(cls), int(k), dev.ptr(out_bits),
                                               nz, ny, nx, dev.stream_ptr()))

    PACKN = 16

    def packn(self, cls, k0, nk, out_bits, vol_words, shape):
        """Seed bits of the radii k0 .. k0 + nk - 1 from one read of the class map (psb200_lt_packn)."""
        nz, ny, nx = shape
        _lib.check(self.ctx.lib.psb200_lt_packn(self.ctx.handle, dev.ptr(cls), int(k0), int(nk), dev.ptr(out_bits),
                                                int(vol_words), nz, ny, nx, dev.stream_ptr()))

    def wmask(self, idx, written, shape):
        nz, ny, nx = shape
        _lib.check(self.ctx.lib.psb200_lt_wmask(self.ctx.handle, dev.ptr(idx), dev.ptr(written), nz, ny, nx,
                                                dev.stream_ptr()))

    def bitball(self, seedbits, nz_src, z_off, written, idx, k, T, shape):
        nz, ny, nx = shape
        _lib.check(self.ctx.lib.psb200_lt_bitball(self.ctx.handle, dev.ptr(seedbits), int(nz_src), int(z_off),
                                                  dev.ptr(written), dev.ptr(idx), int(k), int(T), nz, ny, nx,
                                                  dev.stream_ptr()))

    def expand(self, idx, lut):
        out = self.empty(idx.numel(), self.torch.float64)
        dev.expand_idx(self.ctx, idx, lut, out)
        return out

    def expand_to_host(self, idx, lut, shape, world):
        """float64 slab in (page-locked) host memory; the host threads of the box are shared by the
        ranks, so each rank widens its share of index bytes with cpu_count/world of them."""
        import os
        threads = max(1, (os.cpu_count() or 1) // max(1, world))
        return dev.expand_idx_to_host(self.ctx, idx, lut, shape, nthreads=threads)

    # -- access-limited flooding (slab-local union-find + face flags, psb200_uf_*)
    def uf_begin(self, cls, inlets_u8, shape, z0, nz_global):
        nz, ny, nx = shape
        st = UfState()
        st.cls, st.shape, st.z0, st.nzg = cls, shape, int(z0), int(nz_global)
        st.inlets = inlets_u8
        st.mode = _lib.INLETS_FACES if inlets_u8 is None else _lib.INLETS_MASK
        st.rcls = self.empty(nz * ny * nx, self.torch.uint8)
        st.parent = self.empty(nz * ny * nx + 1, self.torch.int32)
        st.flags = self.zeros(2, self.torch.int32)            # [changed, any marked]
        st.rec = None
        if self.uf_records:
            # row-rooted forest + link records of all radii + join times, built once (csrc/flood_kernels.cuh); the store
            # is the state of the loop, so it is its own allocation rather than the context's shared scratch
            st.rec = self.empty(self.ctx.lib.psb200_uf_records_bytes(self.ctx.handle, nz, ny, nx), self.torch.uint8)
            _lib.check(self.ctx.lib.psb200_uf_begin_records(self.ctx.handle, dev.ptr(cls), dev.ptr(st.parent),
                                                            dev.ptr(st.inlets), st.mode, 3, nz, ny, nx, st.z0, st.nzg,
                                                            dev.ptr(st.rec), st.rec.numel(), dev.stream_ptr()))
            return st
        _lib.check(self.ctx.lib.psb200_uf_begin(self.ctx.handle, dev.ptr(cls), dev.ptr(st.rcls), dev.ptr(st.parent),
                                                dev.ptr(st.inlets), st.mode, 3, nz, ny, nx, st.z0, st.nzg,
                                                dev.stream_ptr()))
        return st

    def uf_activate(self, st, klo, khi):
        nz, ny, nx = st.shape
        if st.rec is not None:
            _lib.check(self.ctx.lib.psb200_uf_activate_records(self.ctx.handle, dev.ptr(st.parent), int(klo), int(khi),
                                                               nz, ny, nx, dev.ptr(st.rec), st.rec.numel(),
                                                               dev.stream_ptr()))
            return
        ws = self.ctx.workspace(self.ctx.lib.psb200_uf_workspace_bytes(self.ctx.handle, nz, ny, nx))
        _lib.check(self.ctx.lib.psb200_uf_activate(self.ctx.handle, dev.ptr(st.parent), dev.ptr(st.cls),
                                                   dev.ptr(st.inlets), st.mode, 3, int(klo), int(khi), nz, ny, nx,
                                                   st.z0, st.nzg, dev.ptr(ws), ws.numel(), dev.stream_ptr()))

    def uf_face(self, st, k, zplane):
        nz, ny, nx = st.shape
        out = self.empty(ny * nx, self.torch.uint8)
        if st.rec is not None:
            _lib.check(self.ctx.lib.psb200_uf_face_records(self.ctx.handle, dev.ptr(st.parent), dev.ptr(st.cls),
                                                           dev.ptr(st.inlets), st.mode, 3, int(k), int(zplane), dev.ptr(out),
                                                           nz, ny, nx, st.z0, st.nzg, dev.ptr(st.rec), st.rec.numel(),
                                                           dev.stream_ptr()))
            return out
        _lib.check(self.ctx.lib.psb200_uf_face(self.ctx.handle, dev.ptr(st.parent), dev.ptr(st.cls), dev.ptr(st.inlets),
                                               st.mode, 3, int(k), int(zplane), dev.ptr(out), nz, ny, nx, st.z0,
                                               st.nzg, dev.stream_ptr()))
        return out

    def uf_inject(self, st, k, zplane, nb_flags):
        nz, ny, nx = st.shape
        if st.rec is not None:
            _lib.check(self.ctx.lib.psb200_uf_inject_records(self.ctx.handle, dev.ptr(st.parent), dev.ptr(st.cls),
                                                             dev.ptr(st.inlets), st.mode, 3, int(k), int(zplane),
                                                             dev.ptr(nb_flags), dev.ptr(st.flags), nz, ny, nx, st.z0,
                                                             st.nzg, dev.ptr(st.rec), st.rec.numel(), dev.stream_ptr()))
            return
        _lib.check(self.ctx.lib.psb200_uf_inject(self.ctx.handle, dev.ptr(st.parent), dev.ptr(st.cls),
                                                 dev.ptr(st.inlets), st.mode, 3, int(k), int(zplane),
                                                 dev.ptr(nb_flags), dev.ptr(st.flags), nz, ny, nx, st.z0, st.nzg,
                                                 dev.stream_ptr()))

    def uf_changed(self, st):
        """1 if an inject since the last call connected something new (reads and clears the flag)."""
        c = int(st.flags[0].item())
        if c:
            st.flags[0] = 0
        return c

    def uf_mark(self, st, k):
        _lib.check(self.ctx.lib.psb200_uf_mark(self.ctx.handle, dev.ptr(st.parent), dev.ptr(st.cls), dev.ptr(st.rcls),
                                               int(k), ctypes.c_void_p(st.flags.data_ptr() + 4), st.cls.numel(),
                                               dev.stream_ptr()))


    def uf_settle(self, st, k):
        """After the exchanges of radius index k.  Job lists: mark the seeds reached at k.  Records: nothing -- the
        join times hold that information until uf_resolve."""
        if st.rec is None:
            self.uf_mark(st, k)

    def uf_resolve(self, st):
        """After the last radius index: rcls = first index at which a voxel is a seed connected to the inlets."""
        if st.rec is not None:
            nz, ny, nx = st.shape
            _lib.check(self.ctx.lib.psb200_uf_resolve_records(self.ctx.handle, dev.ptr(st.parent), dev.ptr(st.cls),
                                                              dev.ptr(st.rcls), nz, ny, nx, dev.ptr(st.rec),
                                                              st.rec.numel(), dev.stream_ptr()))
            st.rec = None              # the store is dead from here on


class UfState:
    """Buffers of one slab's union-find across the radius loop (owned by the backend).

    cls / rcls : class map in, reached-class map out (uint8, flat)      parent : nz*ny*nx + 1 node links (int32)
    inlets, mode : inlet mask (or None) and PSB200_INLETS_*             shape, z0, nzg : slab geometry in the volume
    flags : two device ints [changed by an inject, anything marked]    rec : record store of psb200_uf_*_records, or None
    """
    __slots__ = ("cls", "rcls", "parent", "inlets", "mode", "shape", "z0", "nzg", "flags", "rec")


class ShardedVolume:
    """Driver of the z-slab sharded hot path.  `shape` is the GLOBAL (nz, ny, nx)."""

    def __init__(self, shape, ctx=None, backend=None, group=None):
        import torch
        dist = _dist()
        self.torch = torch
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if len(shape) != 3:
            raise ValueError("ShardedVolume shards 3-D volumes")
        self.shape = tuple(int(s) for s in shape)
        nz, ny, nx = self.shape
        if nz < self.world or ny < self.world:
            raise ValueError(f"volume {self.shape} is too small for {self.world} ranks")
        self.zcounts = split_counts(nz, self.world)
        self.zstarts = [sum(self.zcounts[:r]) for r in range(self.world)]
        self.ysplit, self.ycounts = ceil_split_counts(ny, self.world)
        if min(self.ycounts) == 0:
            raise ValueError(f"ny={ny} cannot be split into {self.world} non-empty pencil ranges")
        self.nzl = self.zcounts[self.rank]
        self.nyl = self.ycounts[self.rank]
        self.backend = backend if backend is not None else CudaBackend(ctx if ctx is not None else _lib.context())

    # ------------------------------------------------------------------ geometry helpers
    @property
    def local_shape(self):
        return (self.nzl, self.shape[1], self.shape[2])

    def local_slice(self):
        z0 = self.zstarts[self.rank]
        return slice(z0, z0 + self.nzl)

    # ---------------------------------------------------------------------- collectives
    def _all_to_all(self, out, inp, out_counts, in_counts):
        if self.world == 1:
            out.copy_(inp)
            return
        _dist().all_to_all_single(out, inp, [int(c) for c in out_counts], [int(c) for c in in_counts],
                                  group=self.group)

    def _allreduce_max(self, value):
        if self.world == 1:
            return int(value)
        t = self.torch.tensor([int(value)], dtype=self.torch.int64, device=self._comm_device())
        _dist().all_reduce(t, op=_dist().ReduceOp.MAX, group=self.group)
        return int(t.item())

    def _comm_device(self):
        return getattr(self.backend, "device", "cpu")

    def exchange_halo(self, planes_lo_send, planes_hi_send, lo_recv, hi_recv):
        """Send my first planes to rank-1 (its hi halo) and my last planes to rank+1 (its lo
        halo); receive my own halos.  Any argument may be None (first / last rank, W = 0)."""
        dist = _dist()
        ops = []
        if self.rank > 0 and planes_lo_send is not None:
            ops.append(dist.P2POp(dist.isend, planes_lo_send, self.rank - 1, self.group))
        if self.rank < self.world - 1 and planes_hi_send is not None:
            ops.append(dist.P2POp(dist.isend, planes_hi_send, self.rank + 1, self.group))
        if self.rank > 0 and lo_recv is not None:
            ops.append(dist.P2POp(dist.irecv, lo_recv, self.rank - 1, self.group))
        if self.rank < self.world - 1 and hi_recv is not None:
            ops.append(dist.P2POp(dist.irecv, hi_recv, self.rank + 1, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def _halo_depths(self, W):
        """Planes of halo below / above my slab for reach W (limited by the neighbour's slab)."""
        nlo = 0 if self.rank == 0 else W
        nhi = 0 if self.rank == self.world - 1 else W
        if W > min(self.zcounts):
            raise ValueError(f"radius reach {W} planes exceeds the thinnest slab ({min(self.zcounts)} planes); "
                             f"use fewer ranks for this volume")
        return nlo, nhi

    # --------------------------------------------------------------------------- inputs
    def blobs(self, porosity=0.5, blobiness=1, seed=0):
        """This rank's slab (uint8 0/1 device tensor) of ONE global `ps.generators.blobs(shape, porosity,
        blobiness)` image (generators/_imgen.py:1023-1051).  The noise is a function of the global voxel index
        (Philox keyed by `seed`), so each rank draws its own planes plus the reach of the z filter: no data
        exchange, and the image is the same for every number of ranks.  Only the per-plane sums of the
        statistics are gathered (fixed order)."""
        from . import generators as gen
        be, torch = self.backend, self.torch
        ctx = be.ctx
        nz, ny, nx = self.shape
        _, sigma = gen._prologue(self.shape, blobiness, 1)
        z0, nzl, plane = self.zstarts[self.rank], self.nzl, ny * nx
        rz = gen.gaussian_kernel(sigma[0])[0] if float(sigma[0]) > 1e-15 else 0
        if rz >= nz:
            raise ValueError("blobs: the z filter is longer than the volume")
        # global planes the z correlation of my slab touches (reflection at the global ends folds back inside)
        lo, hi = max(0, z0 - rz), min(nz, z0 + nzl + rz)
        if z0 - rz < 0:
            hi = max(hi, min(nz, rz - z0))
        if z0 + nzl + rz > nz:
            lo = min(lo, max(0, 2 * nz - (z0 + nzl + rz)))
        a = gen.philox_noise(ctx, (hi - lo) * plane, seed, first=lo * plane)
        b = torch.empty(nzl * plane, dtype=torch.float64, device=a.device)
        if rz > 0 or float(sigma[0]) > 1e-15:
            gen.gauss_axis(ctx, a, b, 0, sigma[0], (nzl, ny, nx), z_out0=z0, z_in0=lo, nz_in=hi - lo, nz_glob=nz)
        else:
            b.copy_(a[(z0 - lo) * plane:(z0 - lo + nzl) * plane])
        del a                                   # (the noise planes: at 2048^3 on one GPU they are 64 GiB)
        a = torch.empty_like(b)
        for axis in (1, 2):
            if float(sigma[axis]) > 1e-15:
                gen.gauss_axis(ctx, b, a, axis, sigma[axis], (nzl, ny, nx))
                a, b = b, a
        f = b
        del a
        s, flo, fhi = gen.field_stats(ctx, f, nzl, plane)
        n = nz * ny * nx
        mean = gen.combine_sum(self._gather_planes(s)) / n
        sq = gen._stat(ctx, f, nzl, plane, mean, 1)
        sd = float(np.sqrt(gen.combine_sum(self._gather_planes(sq)) / n))
        flo = -self._allreduce_max_f64(-flo)
        fhi = self._allreduce_max_f64(fhi)
        return gen.finish(ctx, f, nzl * plane, mean, sd, flo, fhi, porosity)

    def _gather_planes(self, part):
        """Per-plane partial sums of every rank, in global plane order (numpy [nz][chunks])."""
        if self.world == 1:
            return part
        out = [None] * self.world
        _dist().all_gather_object(out, np.asarray(part), group=self.group)
        return np.concatenate(out, axis=0)

    def _allreduce_max_f64(self, value):
        if self.world == 1:
            return float(value)
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self._comm_device())
        _dist().all_reduce(t, op=_dist().ReduceOp.MAX, group=self.group)
        return float(t.item())

    # ------------------------------------------------------------------------------ EDT
    # Fast path of the sharded EDT: every rank receives EDT_HALO input planes from each z-neighbour (1 byte per
    # voxel, 1/4 of one plane of distances) and runs the one-GPU EDT on its extended slab.  A voxel of the own
    # slab is at least EDT_HALO + 1 planes away from anything outside the extended slab, so every squared
    # distance below (EDT_HALO + 1)^2 is exact; if the global maximum stays below that bound -- porous media: the
    # largest pore radius is a few tens of voxels -- the result is the exact EDT and neither all-to-all
    # transpose is needed.  Otherwise (or with `edt_halo = 0`) the slab -> pencil all-to-all path runs, which is
    # exact for any input.
    EDT_HALO = 64

    def _edt_halo(self, local_u8):
        """(own-slab squared distances as flat int32, global max d2) through the input-halo path, or None when
        the bound does not hold (or the slabs are thinner than the halo)."""
        torch, be = self.torch, self.backend
        nz, ny, nx = self.shape
        nzl, P, H = self.nzl, self.world, int(getattr(self, "edt_halo", self.EDT_HALO))
        if P == 1 or H <= 0 or H > min(self.zcounts) or not hasattr(be, "edt_ext"):
            return None
        plane = ny * nx
        lo = H if self.rank > 0 else 0
        hi = H if self.rank < P - 1 else 0
        flat = local_u8.reshape(-1)
        ext = be.empty((lo + nzl + hi) * plane, torch.uint8)
        ext[lo * plane:(lo + nzl) * plane] = flat
        if hasattr(be, "mask_pack") and (H * plane) % 8 == 0 and (lo * plane) % 8 == 0 and ((lo + nzl) * plane) % 8 == 0:
            # the halo planes travel as bits: an eighth of the bytes
            nb = H * plane // 8
            r_lo = be.empty(nb, torch.uint8) if lo else None
            r_hi = be.empty(nb, torch.uint8) if hi else None
            self.exchange_halo(be.mask_pack(flat[:H * plane]) if self.rank > 0 else None,
                               be.mask_pack(flat[(nzl - H) * plane:]) if self.rank < P - 1 else None, r_lo, r_hi)
            if lo:
                be.mask_unpack(r_lo, ext[:lo * plane])
            if hi:
                be.mask_unpack(r_hi, ext[(lo + nzl) * plane:])
        else:
            self.exchange_halo(flat[:H * plane] if self.rank > 0 else None,
                               flat[(nzl - H) * plane:] if self.rank < P - 1 else None,
                               ext[:lo * plane] if lo else None, ext[(lo + nzl) * plane:] if hi else None)
        own, lmax = be.edt_ext(ext, (lo + nzl + hi, ny, nx), lo, nzl)
        del ext
        gmax = self._allreduce_max(lmax)
        if gmax >= (H + 1) * (H + 1):          # includes INF (a slab without background): not provably exact
            return None
        return own, gmax

    def edt_sq(self, local_u8):
        """Local slab (uint8, flat or [nzl][ny][nx]) -> (uint32 squared distances of the slab as a
        flat int32 tensor, global max d2)."""
        torch, be = self.torch, self.backend
        nz, ny, nx = self.shape
        nzl, nyl, P = self.nzl, self.nyl, self.world
        fast = self._edt_halo(local_u8)
        self.edt_path = "halo" if fast is not None else "all-to-all"
        if fast is not None:
            return fast
        d2p, lmax = self._edt_pencils(local_u8)
        if P == 1:
            return d2p, lmax
        return self._pencils_to_slab(d2p, torch.int32), self._allreduce_max(lmax)

    def _edt_pencils(self, local_u8):
        """x / y passes on the slab, slab -> pencil all-to-all, z pass on the pencils.  Returns the
        squared distances in pencil layout [nz][nyl][nx] (the slab itself for one rank) and the
        local max."""
        torch, be = self.torch, self.backend
        nz, ny, nx = self.shape
        nzl, nyl, P = self.nzl, self.nyl, self.world
        h_send = be.edt_xy(local_u8.reshape(-1), (nzl, ny, nx), self.ysplit if P > 1 else 0)
        if P == 1:
            return be.edt_z(h_send, (nz, ny, nx))
        # slab -> pencil: block for dest d is [nzl][ycounts[d]][nx]; I receive [zcounts[s]][nyl][nx] from s
        pencil = be.empty(nz * nyl * nx, torch.int32)
        self._all_to_all(pencil, h_send, [self.zcounts[s] * nyl * nx for s in range(P)],
                         [nzl * self.ycounts[d] * nx for d in range(P)])
        del h_send
        return be.edt_z(pencil, (nz, nyl, nx))

    def _pencils_to_slab(self, vals_p, dtype):
        """pencil -> slab: dest d gets my rows for its planes (a contiguous z range of the pencil).
        Works for any element type (uint32 distances, uint8 classes)."""
        be = self.backend
        nz, ny, nx = self.shape
        nzl, nyl, P = self.nzl, self.nyl, self.world
        recv = be.empty(nzl * ny * nx, dtype)
        self._all_to_all(recv, vals_p, [nzl * self.ycounts[s] * nx for s in range(P)],
                         [self.zcounts[d] * nyl * nx for d in range(P)])
        slab = be.empty(nzl * ny * nx, dtype).view(nzl, ny, nx)
        off = 0
        for s in range(P):
            cnt = nzl * self.ycounts[s] * nx
            y0 = s * self.ysplit
            slab[:, y0:y0 + self.ycounts[s], :] = recv[off:off + cnt].view(nzl, self.ycounts[s], nx)
            off += cnt
        return slab.reshape(-1)

    def edt(self, local_im):
        """float32 distances of the local slab (edt.edt semantics on the global volume)."""
        torch = self.torch
        d2, _ = self.edt_sq(self.backend.to_u8(local_im))
        u = d2.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        out = torch.sqrt(u.to(torch.float32))                      # d2 < 2^24: exact float32, IEEE sqrt
        out = torch.where(u == host.INF_U32, torch.full_like(out, float("inf")), out)
        return out.view(self.nzl, self.shape[1], self.shape[2])

    # --------------------------------------------------------------------- radius loop
    def local_thickness(self, local_im, sizes=25, to_host=False, as_index=False):
        """ps.filters.local_thickness of the GLOBAL volume; returns this rank's slab (float64):
        a device tensor, or with `to_host` a numpy array (the reference's return type).  `as_index`: the slab in
        index form instead, (uint8 device tensor, float64 table): map = table[index] (sizemap.IndexMap layout)."""
        return self._porosimetry(local_im, sizes, access_limited=False, to_host=to_host, as_index=as_index)

    def porosimetry(self, local_im, sizes=25, inlets=None, access_limited=True, to_host=False, as_index=False):
        """ps.filters.porosimetry of the GLOBAL volume (F:1032-1212); returns this rank's slab.
        `inlets`: None = all faces of the global volume (F:1128-1129), else this rank's slab
        [nzl][ny][nx] of the global inlet mask."""
        return self._porosimetry(local_im, sizes, access_limited=access_limited, inlets=inlets, to_host=to_host,
                                 as_index=as_index)

    def _flood_exchange(self, st, k):
        """Propagate inlet connectivity through the slab faces until no rank learns anything new
        (trim_disconnected_blobs on the global volume, F:1252-1270).  Returns the sweep count."""
        be = self.backend
        nzl, ny, nx = self.local_shape
        if self.world == 1:
            return 0
        sweeps = 0
        while True:
            sweeps += 1
            lo_send = be.uf_face(st, k, 0) if self.rank > 0 else None
            hi_send = be.uf_face(st, k, nzl - 1) if self.rank < self.world - 1 else None
            lo_recv = be.empty(ny * nx, self.torch.uint8) if self.rank > 0 else None
            hi_recv = be.empty(ny * nx, self.torch.uint8) if self.rank < self.world - 1 else None
            self.exchange_halo(lo_send, hi_send, lo_recv, hi_recv)
            if lo_recv is not None:
                be.uf_inject(st, k, 0, lo_recv)
            if hi_recv is not None:
                be.uf_inject(st, k, nzl - 1, hi_recv)
            if not self._allreduce_max(be.uf_changed(st)):
                return sweeps

    def _porosimetry(self, local_im, sizes, access_limited, inlets=None, to_host=False, as_index=False):
        torch, be = self.torch, self.backend
        nz, ny, nx = self.shape
        nzl = self.nzl
        lshape = (nzl, ny, nx)
        # the radius loop only needs the class of every voxel, so the distances are classified on the
        # pencils and ONE byte per voxel travels back to the slabs instead of four
        local_u8 = be.to_u8(local_im, positive=True)                           # F:1126 edt(im > 0)
        fast = self._edt_halo(local_u8)
        self.edt_path = "halo" if fast is not None else "all-to-all"
        if fast is not None:
            d2p, max_d2 = fast               # already in slab layout
        else:
            d2p, lmax = self._edt_pencils(local_u8)
            max_d2 = self._allreduce_max(lmax)
        del local_u8
        self.last_max_d2 = max_d2
        radii = host.reference_sizes(sizes, max_d2)
        if max_d2 == host.INF_U32:
            from .filters import _result_for_no_background
            res = _result_for_no_background(lshape, radii)
            return res if to_host else torch.from_numpy(res).to(self._comm_device())
        T, R = host.effective_thresholds(radii, max_d2)
        if len(T) > _lib.MAX_THRESHOLDS:
            raise NotImplementedError("sharded path supports up to 253 effective radii per call")
        if len(T) and self.world > 1:
            self._halo_depths(host.isqrt(int(T[0]) - 1))     # the deepest reach must fit the thinnest slab: fail before any exchange
        n = nzl * ny * nx
        idx = be.zeros(n, torch.uint8)
        if len(T) == 0:
            return be.expand(idx, np.array([0.0])).view(*lshape)
        cls = be.classify(d2p, T)
        del d2p
        if self.world > 1 and fast is None:
            cls = self._pencils_to_slab(cls, torch.uint8)
        st = None
        if access_limited:
            inl = None
            if inlets is not None:
                if tuple(np.shape(inlets)) != lshape:
                    raise Exception("inlets not valid, refer to docstring for info")
                inl = be.to_u8(inlets)
            st = be.uf_begin(cls, inl, lshape, self.zstarts[self.rank], nz)
            # F:1181-1183 for every radius before the first dilation: the seed sets are nested, so the union-find only
            # links the voxels that became seeds at each radius, and the dilation of radius k only asks whether a
            # voxel's FIRST radius as a reached seed is k (or <= k) -- the final map answers that for every k
            self.flood_sweeps = []
            for k in range(len(T)):
                be.uf_activate(st, k - 1, k)
                self.flood_sweeps.append(self._flood_exchange(st, k))
                be.uf_settle(st, k)
            be.uf_resolve(st)
            cls = st.rcls
            st.parent = None
        written = None
        packed, packed_bits = None, None
        # The seed bits of every bit-path radius are a function of the (reached-)class map alone, which is final
        # here, so the class bytes of the deepest bit-path halo travel ONCE and each rank packs its
        # extended slab itself -- one halo exchange instead of one per bit-path radius.
        cls_ext, ext_lo, ext_hi = None, 0, 0
        if self.world > 1:
            wb = [host.isqrt(int(Tk) - 1) for Tk in T if be.bit_ok(lshape, int(Tk))]
            if wb and max(wb) > 0:
                ext_lo, ext_hi = self._halo_depths(max(wb))
                plane_b = ny * nx
                cls_ext = be.empty((ext_lo + nzl + ext_hi) * plane_b, torch.uint8)
                cls_ext[ext_lo * plane_b:(ext_lo + nzl) * plane_b] = cls
                self.exchange_halo(cls[:ext_lo * plane_b] if ext_lo else None,
                                   cls[(nzl - ext_hi) * plane_b:] if ext_hi else None,
                                   cls_ext[:ext_lo * plane_b] if ext_lo else None,
                                   cls_ext[(ext_lo + nzl) * plane_b:] if ext_hi else None)
        for k, Tk in enumerate(T):
            Tk = int(Tk)
            W = host.isqrt(Tk - 1)
            nlo, nhi = self._halo_depths(W)
            if be.bit_ok(lshape, Tk):
                nw = nx // 32
                plane = ny * nw
                if written is None:
                    written = be.zeros(n // 32, torch.int32)
                    if k > 0:
                        be.wmask(idx, written, lshape)
                if cls_ext is not None:
                    nze = ext_lo + nzl + ext_hi
                    if packed is None or not (packed[0] <= k < packed[1]):
                        # every later radius is a bit radius too (thresholds descend): pack up to PACKN of them
                        # from one read of the extended class map
                        nk = min(getattr(be, "PACKN", 1), len(T) - k)
                        vol_words = (nze * plane + 63) & ~63
                        packed_bits = be.empty(nk * vol_words, torch.int32)
                        if nk > 1:
                            be.packn(cls_ext, k, nk, packed_bits, vol_words, (nze, ny, nx))
                        else:
                            be.pack(cls_ext, k, packed_bits[:nze * plane], (nze, ny, nx))
                        packed = (k, k + nk, vol_words)
                    off = (k - packed[0]) * packed[2]
                    be.bitball(packed_bits[off:off + nze * plane], nze, ext_lo, written, idx, k, Tk, lshape)
                    continue
                ext = be.empty((nlo + nzl + nhi) * plane, torch.int32)
                mine = ext[nlo * plane:(nlo + nzl) * plane]
                be.pack(cls, k, mine, lshape)
                self.exchange_halo(mine[:nlo * plane] if self.rank > 0 and W else None,
                                   mine[(nzl - nhi) * plane:] if nhi else None,
                                   ext[:nlo * plane] if nlo else None,
                                   ext[(nlo + nzl) * plane:] if nhi else None)
                be.bitball(ext, nlo + nzl + nhi, nlo, written, idx, k, Tk, lshape)
                del ext
            else:
                reach = be.lt_xy(cls, k, Tk, lshape)
                plane = ny * nx
                # a neighbour's cone sweep needs one number per column from this slab: the cone value arriving at the
                # shared face (cones are shorter than the thinnest slab, so nothing passes through a whole slab) --
                # one plane per direction and radius instead of W
                m_lo = be.empty(plane, torch.uint8) if nlo else None
                m_hi = be.empty(plane, torch.uint8) if nhi else None
                self.exchange_halo(be.halo_cone(reach, lshape, W, 0) if self.rank > 0 and W else None,
                                   be.halo_cone(reach, lshape, W, 1) if nhi else None, m_lo, m_hi)
                be.lt_z(reach, m_lo, m_hi, idx, k, Tk, lshape)
                del reach
        lut = np.concatenate([[0.0], R])
        if as_index:
            return idx, lut
        if to_host and hasattr(be, "expand_to_host"):
            return be.expand_to_host(idx, lut, lshape, self.world)
        return be.expand(idx, lut).view(*lshape)
