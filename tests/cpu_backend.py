"""TEST INFRASTRUCTURE: numpy/oracle implementation of the step primitives that
porespy_b200.sharded.ShardedVolume drives, on CPU torch tensors, so the partitioning,
all-to-all and halo-exchange logic runs under gloo without a GPU.  Each method mirrors the
contract of one libpsb200.so entry point (include/psb200.h); the product never imports this."""
import numpy as np
import torch

from oracle import cpu as oc
from tests import model_fast as mf

INF = 0xFFFFFFFF


class CpuBackend:
    device = "cpu"

    def __init__(self, bit_tmax=200):
        self.bit_tmax = bit_tmax

    def empty(self, n, dtype):
        return torch.empty(int(n), dtype=dtype)

    def zeros(self, n, dtype):
        return torch.zeros(int(n), dtype=dtype)

    def to_u8(self, arr, positive=False):
        if isinstance(arr, torch.Tensor):
            arr = arr.numpy()
        arr = np.asarray(arr)
        mask = (arr > 0) if positive else (arr != 0)
        return torch.from_numpy(np.ascontiguousarray(mask).view(np.uint8).copy()).reshape(-1)

    def bit_ok(self, shape, T):
        return shape[2] % 32 == 0 and T <= self.bit_tmax

    @staticmethod
    def _u32(t):
        return t.numpy().view(np.uint32)

    # psb200_edt_xy_u8: exact 2-D squared distances per plane, optionally in all-to-all send layout
    def edt_xy(self, im_u8, shape, ysplit):
        nz, ny, nx = shape
        im = im_u8.numpy().reshape(shape)
        h = np.empty(shape, dtype=np.uint32)
        for z in range(nz):
            h[z] = oc.edt_sq(im[z])
        if ysplit:
            blocks = [h[:, y0:y0 + ysplit, :].reshape(-1) for y0 in range(0, ny, ysplit)]
            h = np.concatenate(blocks)
        return torch.from_numpy(h.reshape(-1).view(np.int32).copy())

    def edt_ext(self, ext_u8, shape_ext, lo, nzl):
        nze, ny, nx = shape_ext
        d2 = oc.edt_sq(ext_u8.numpy().reshape(shape_ext))[lo:lo + nzl]
        return torch.from_numpy(d2.reshape(-1).view(np.int32).copy()), int(d2.max()) if d2.size else 0

    # psb200_edt_z_u32: out[z] = min_z' h[z'] + (z - z')^2
    def edt_z(self, h, shape):
        nz, ny, nx = shape
        hv = self._u32(h).reshape(shape).astype(np.int64)
        big = np.int64(1) << 40
        hv = np.where(hv == INF, big, hv)
        out = np.full(shape, big, dtype=np.int64)
        zs = np.arange(nz)
        for zp in range(nz):
            cand = hv[zp][None, :, :] + ((zs - zp) ** 2)[:, None, None]
            np.minimum(out, cand, out=out)
        out = np.where(out >= big, INF, out).astype(np.uint32)
        mx = int(out.max()) if out.size else 0
        return torch.from_numpy(out.reshape(-1).view(np.int32).copy()), mx

    def classify(self, d2, T):
        return torch.from_numpy(mf.classify(self._u32(d2), np.asarray(T)).copy())

    def lt_xy(self, cls, k, T, shape):
        return torch.from_numpy(mf.reach_map(cls.numpy().reshape(shape), k, int(T)).reshape(-1).copy())

    def mask_pack(self, src_u8):
        return torch.from_numpy(np.packbits(src_u8.numpy() != 0, bitorder="little"))

    def mask_unpack(self, bits, dst_u8):
        dst_u8.numpy()[:] = np.unpackbits(bits.numpy(), bitorder="little")[:dst_u8.numel()]

    def halo_cone(self, reach, shape, depth, side):
        """psb200_lt_halo_cone: max_j (reach[plane j from the face] - j), j < depth."""
        nz, ny, nx = shape
        r = reach.numpy().reshape(shape).astype(np.int64)
        out = np.zeros((ny, nx), dtype=np.int64)
        for j in range(min(int(depth), nz)):
            out = np.maximum(out, r[nz - 1 - j if side else j] - j)
        return torch.from_numpy(out.astype(np.uint8).reshape(-1))

    def lt_z(self, reach, m_lo, m_hi, idx, k, T, shape):
        nz, ny, nx = shape
        parts, nlo = [], 0
        if m_lo is not None and m_lo.numel():
            parts.append(m_lo.numpy().reshape(-1, ny, nx))
            nlo = parts[0].shape[0]
        parts.append(reach.numpy().reshape(shape))
        if m_hi is not None and m_hi.numel():
            parts.append(m_hi.numpy().reshape(-1, ny, nx))
        fill = mf.cone_fill(np.concatenate(parts, axis=0))[nlo:nlo + nz]
        iv = idx.numpy().reshape(shape)
        iv[(iv == 0) & fill] = k + 1

    def pack(self, cls, k, out_bits, shape):
        bits = np.packbits(cls.numpy() <= k, bitorder="little")
        out_bits.numpy().view(np.uint8)[:] = bits

    PACKN = 3

    def packn(self, cls, k0, nk, out_bits, vol_words, shape):
        n = cls.numel() // 32
        for i in range(nk):
            self.pack(cls, k0 + i, out_bits[i * vol_words:i * vol_words + n], shape)

    def wmask(self, idx, written, shape):
        written.numpy().view(np.uint8)[:] = np.packbits(idx.numpy() != 0, bitorder="little")

    def bitball(self, seedbits, nz_src, z_off, written, idx, k, T, shape):
        nz, ny, nx = shape
        seeds = np.unpackbits(seedbits.numpy().view(np.uint8), bitorder="little").astype(bool).reshape(nz_src, ny, nx)
        if seeds.any():
            fill = (oc.edt_sq(~seeds) < T)[z_off:z_off + nz]
        else:
            fill = np.zeros(shape, dtype=bool)
        wr = np.unpackbits(written.numpy().view(np.uint8), bitorder="little").astype(bool).reshape(shape)
        iv = idx.numpy().reshape(shape)
        new = fill & ~wr
        assert not (iv[new] != 0).any(), "written mask out of sync with idx"
        iv[new] = k + 1
        written.numpy().view(np.uint8)[:] = np.packbits(wr | fill, bitorder="little")

    def expand(self, idx, lut):
        return torch.from_numpy(np.asarray(lut, dtype=np.float64)[idx.numpy()].copy())

    # psb200_uf_*: slab-local connectivity to the inlets + links injected through the slab faces
    def uf_begin(self, cls, inlets_u8, shape, z0, nz_global):
        import types
        nz, ny, nx = shape
        st = types.SimpleNamespace()
        st.cls = cls.numpy().reshape(shape)
        if inlets_u8 is None:          # get_border(global shape, 'faces') restricted to this slab
            m = np.zeros(shape, dtype=bool)
            zg = np.arange(z0, z0 + nz)
            m[(zg == 0) | (zg == nz_global - 1)] = True
            m[:, [0, -1], :] = True
            m[:, :, [0, -1]] = True
            st.inlet = m
        else:
            st.inlet = inlets_u8.numpy().reshape(shape) != 0
        st.injected = np.zeros(shape, dtype=bool)
        st.rcls = torch.from_numpy(np.where(st.cls == mf.CLS_BG, mf.CLS_BG, mf.CLS_NEVER).astype(np.uint8).reshape(-1))
        st.shape, st.changed = shape, 0
        return st

    def _connected(self, st, k):
        import scipy.ndimage as spim
        nodes = st.inlet | (st.cls <= k)
        lab = spim.label(nodes, structure=spim.generate_binary_structure(3, 1))[0]
        keep = np.unique(lab[(st.inlet | st.injected) & nodes])
        return nodes, np.isin(lab, keep[keep > 0])

    def uf_activate(self, st, klo, khi):
        pass

    def uf_face(self, st, k, zplane):
        return torch.from_numpy(self._connected(st, k)[1][zplane].astype(np.uint8).reshape(-1).copy())

    def uf_inject(self, st, k, zplane, nb_flags):
        nodes, conn = self._connected(st, k)
        ny, nx = st.shape[1:]
        new = nodes[zplane] & (nb_flags.numpy().reshape(ny, nx) != 0) & ~conn[zplane]
        if new.any():
            st.injected[zplane] |= new
            st.changed = 1

    def uf_changed(self, st):
        c, st.changed = st.changed, 0
        return c

    def uf_mark(self, st, k):
        conn = self._connected(st, k)[1]
        r = st.rcls.numpy().reshape(st.shape)
        r[(st.cls <= k) & conn & (r == mf.CLS_NEVER)] = k

    # the driver's view of the two GPU flavours: per-radius marking (here) or join times + one resolve pass
    def uf_settle(self, st, k):
        self.uf_mark(st, k)

    def uf_resolve(self, st):
        pass
