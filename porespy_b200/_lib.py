"""ctypes binding of libpsb200.so (the C ABI declared in include/psb200.h).

The product path has no CPU fallback: if the shared library is missing or no CUDA device is
visible, every public entry point raises -- it never routes through `oracle/`.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpsb200.so")
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "psb200.h")

OK = 0
INLETS_NONE, INLETS_FACES, INLETS_MASK = 0, 1, 2
ALGO_FAST, ALGO_GENERIC = 0, 1
FLAG_IDX_PREINIT = 1
FLAG_EXPAND_MERGE = 1
MAX_THRESHOLDS = 253
MAX_DIM = 32767
INF_U32 = 0xFFFFFFFF

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


class Psb200Error(RuntimeError):
    pass


def build(force=False, verbose=False):
    """Compile csrc/psb200.cu for sm_100a into porespy_b200/libpsb200.so (in-tree)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [HEADER]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, "psb200.cu"), "-o", LIB_PATH]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB_PATH


_c = ctypes
_vp, _i64, _i32, _u32, _sz = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_uint32, _c.c_size_t

# name -> (restype, argtypes); mirrors include/psb200.h one to one
SIGNATURES = {
    "psb200_version": (_i32, []),
    "psb200_last_error": (_c.c_char_p, []),
    "psb200_create": (_i32, [_i32, _c.POINTER(_vp)]),
    "psb200_destroy": (_i32, [_vp]),
    "psb200_set_option": (_i32, [_vp, _c.c_char_p, _i64]),
    "psb200_launch_count": (_i64, [_vp]),
    "psb200_profile_kernels": (_i32, []),
    "psb200_profile_name": (_c.c_char_p, [_i32]),
    "psb200_profile_read": (_i32, [_vp, _c.POINTER(_c.c_double), _c.POINTER(_i64)]),
    "psb200_profile_records": (_i32, [_vp, _c.POINTER(_i32), _c.POINTER(_c.c_float), _i32]),
    "psb200_edt_workspace_bytes": (_sz, [_vp, _i64, _i64, _i64]),
    "psb200_edt_sq_u8": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _sz, _vp]),
    "psb200_edt_u8": (_i32, [_vp, _vp, _vp, _i32, _vp, _i64, _i64, _i64, _vp, _sz, _vp]),
    "psb200_edt_u8_zmax": (_i32, [_vp, _vp, _vp, _i32, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _sz, _vp]),
    "psb200_edt_xy_u8": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _sz, _vp]),
    "psb200_edt_z_u32": (_i32, [_vp, _vp, _vp, _i32, _vp, _i64, _i64, _i64, _vp]),
    "psb200_edt_pass": (_i32, [_vp, _i32, _vp, _vp, _i64, _i64, _i64, _vp, _sz, _vp]),
    "psb200_sqrt_f32": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "psb200_max_u32": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "psb200_local_thickness_workspace_bytes": (_sz, [_vp, _i64, _i64, _i64, _i32]),
    "psb200_local_thickness_idx": (_i32, [_vp, _vp, _c.POINTER(_u32), _i32, _vp, _vp, _i32, _i32,
                                          _i64, _i64, _i64, _i32, _vp, _sz, _vp]),
    "psb200_lt_classify": (_i32, [_vp, _vp, _c.POINTER(_u32), _i32, _vp, _i64, _vp]),
    "psb200_lt_xy": (_i32, [_vp, _vp, _i32, _u32, _vp, _i64, _i64, _i64, _vp, _sz, _vp]),
    "psb200_lt_z": (_i32, [_vp, _vp, _vp, _i32, _vp, _i32, _vp, _i32, _u32, _i64, _i64, _i64, _vp]),
    "psb200_lt_pack": (_i32, [_vp, _vp, _i32, _vp, _i64, _i64, _i64, _vp]),
    "psb200_lt_packn": (_i32, [_vp, _vp, _i32, _i32, _vp, _i64, _i64, _i64, _i64, _vp]),
    "psb200_lt_wmask": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "psb200_lt_bitball": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _i32, _u32, _i64, _i64, _i64, _vp]),
    "psb200_expand_idx_f64": (_i32, [_vp, _vp, _c.POINTER(_c.c_double), _i32, _vp, _i64, _i32, _vp]),
    "psb200_expand_idx_f64_to_host": (_i32, [_vp, _vp, _c.POINTER(_c.c_double), _i32, _vp, _i64, _vp, _sz,
                                             _vp, _sz, _i32, _i32, _i32, _vp]),
    "psb200_upload_mask_u8": (_i32, [_vp, _vp, _i64, _vp, _vp, _sz, _vp, _sz, _i32, _vp]),
    "psb200_mark_written": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "psb200_uf_begin": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _i64, _i64, _i64, _i64, _vp]),
    "psb200_mask_pack_u8": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "psb200_mask_unpack_u8": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "psb200_lt_halo_cone": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp]),
    "psb200_uf_workspace_bytes": (_sz, [_vp, _i64, _i64, _i64]),
    "psb200_uf_records_bytes": (_sz, [_vp, _i64, _i64, _i64]),
    "psb200_uf_begin_records": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i64, _i64, _i64, _i64, _i64, _vp, _sz, _vp]),
    "psb200_uf_activate_records": (_i32, [_vp, _vp, _i32, _i32, _i64, _i64, _i64, _vp, _sz, _vp]),
    "psb200_uf_face_records": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i64, _vp, _i64, _i64, _i64, _i64, _i64,
                                      _vp, _sz, _vp]),
    "psb200_uf_inject_records": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i64, _vp, _vp, _i64, _i64, _i64, _i64,
                                        _i64, _vp, _sz, _vp]),
    "psb200_uf_resolve_records": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _sz, _vp]),
    "psb200_uf_activate": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _i64, _i64,
                                  _vp, _sz, _vp]),
    "psb200_uf_face": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i64, _vp, _i64, _i64, _i64, _i64, _i64, _vp]),
    "psb200_uf_inject": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp]),
    "psb200_uf_mark": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp]),
    "psb200_flood_workspace_bytes": (_sz, [_vp, _i64, _i64, _i64]),
    "psb200_noise_philox_f64": (_i32, [_vp, _vp, _i64, _c.c_uint64, _c.c_uint64, _vp]),
    "psb200_gauss_workspace_bytes": (_sz, [_vp, _i32]),
    "psb200_gauss_axis_f64": (_i32, [_vp, _vp, _vp, _i32, _c.POINTER(_c.c_double), _i32, _i64, _i64, _i64, _i64, _i64,
                                     _i64, _i64, _vp, _sz, _vp]),
    "psb200_stats_chunks": (_i32, []),
    "psb200_stats_f64": (_i32, [_vp, _vp, _i64, _i64, _c.c_double, _i32, _vp, _vp]),
    "psb200_blobs_finish": (_i32, [_vp, _vp, _i64, _c.c_double, _c.c_double, _c.c_double, _c.c_double, _c.c_double,
                                   _vp, _vp, _vp]),
    "psb200_hist_idx": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, _vp, _vp]),
    "psb200_expand_lut8": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _vp]),
    "psb200_expand_lut1": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _vp]),
    "psb200_distinct64": (_i32, [_vp, _vp, _i64, _vp, _u32, _vp, _vp]),
    "psb200_index_of64": (_i32, [_vp, _vp, _i64, _vp, _i32, _i32, _vp, _i32, _vp]),
    "psb200_drain_stats": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _c.c_double, _c.c_double, _c.c_double, _i32, _vp, _i32, _vp]),
    "psb200_drain_threshold": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _c.c_double, _c.c_double, _c.c_double, _i32,
                                      _c.c_double, _vp, _vp]),
    "psb200_drain_newly": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "psb200_drain_classify": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _c.c_double, _c.c_double, _c.c_double, _i32,
                                     _vp, _i32, _vp, _vp]),
    "psb200_drain_newly_rcls": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "psb200_flood_classes": (_i32, [_vp, _vp, _vp, _i32, _vp, _i32, _i32, _i64, _i64, _i64, _vp, _sz, _vp]),
    "psb200_drain_paint_workspace_bytes": (_sz, [_vp, _i64, _i64, _i64]),
    "psb200_drain_paint": (_i32, [_vp, _vp, _i32, _vp, _i32, _i64, _i64, _i64, _vp, _sz, _vp]),
    "psb200_set_where_u8": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp]),
    "psb200_set_zero_codes_u8": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _vp]),
    "psb200_flood": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _sz, _vp]),
}

_lib = None
_lock = threading.Lock()


def load():
    """dlopen libpsb200.so and attach the prototypes.  Raises if the library is absent."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise Psb200Error(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                    f"g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def check(rc):
    if rc != OK:
        msg = load().psb200_last_error().decode("utf-8", "replace")
        raise Psb200Error(f"psb200 error {rc}: {msg}")


class Context:
    """One psb200 context per CUDA device (per rank)."""

    def __init__(self, device=0):
        import torch
        if not torch.cuda.is_available():
            raise Psb200Error("porespy_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.lib = load()
        self.device = int(device)
        self.handle = _vp()
        check(self.lib.psb200_create(self.device, ctypes.byref(self.handle)))
        self._ws = None

    def set_algo(self, algo):
        check(self.lib.psb200_set_option(self.handle, b"algo", int(algo)))

    def set_bit_tmax(self, tmax):
        """Thresholds T <= tmax run the bit-parallel dilation kernels (0 = never)."""
        check(self.lib.psb200_set_option(self.handle, b"bit_tmax", int(tmax)))

    def set_bit4(self, on):
        """Bit path: four-words-per-lane dilation kernel for rows of 1024 / 2048 / 4096 voxels (default);
        off = the one- / two-words-per-lane kernels for every shape."""
        check(self.lib.psb200_set_option(self.handle, b"bit4", 1 if on else 0))

    def set_edt16(self, on):
        """EDT y/z passes: run the 16-bit two-voxels-per-instruction kernel first (default); off =
        the uint32 kernel only."""
        check(self.lib.psb200_set_option(self.handle, b"edt16", 1 if on else 0))

    def set_foot(self, foot):
        """Warp footprint of the 16-bit min-plus scans (EDT y/z passes, per-radius y pass): 0 = 64 columns x
        8 rows, 1 = 32 x 16."""
        check(self.lib.psb200_set_option(self.handle, b"foot", int(foot)))

    def set_uf_records(self, on):
        """Flood: row-rooted forest + link records (default on) / the per-voxel job lists."""
        check(self.lib.psb200_set_option(self.handle, b"uf_records", 1 if on else 0))

    def set_yflags(self, on):
        """Byte path: activity flags from the x pass let the y pass skip idle tiles and rows (default on)."""
        check(self.lib.psb200_set_option(self.handle, b"yflags", 1 if on else 0))

    def set_zwide(self, on):
        """z sweeps: 8 columns per thread, 16 planes in flight (default) / 4 columns, 8 planes."""
        check(self.lib.psb200_set_option(self.handle, b"zwide", 1 if on else 0))

    def set_edt_h(self, rows):
        """Halo rows of the 16-bit EDT tiles (default 32)."""
        check(self.lib.psb200_set_option(self.handle, b"edt_h", int(rows)))

    def set_ydirect(self, on):
        """Per-radius y pass: store reach bytes directly from registers (default) / through a shared-memory tile."""
        check(self.lib.psb200_set_option(self.handle, b"ydirect", 1 if on else 0))

    def set_bitquad(self, on):
        """Bit path: two output rows per lane (default) / one output row per lane."""
        check(self.lib.psb200_set_option(self.handle, b"bitquad", 1 if on else 0))

    def set_xbits(self, on):
        """Per-radius x pass from packed seed bits (default) / from the class bytes."""
        check(self.lib.psb200_set_option(self.handle, b"xbits", 1 if on else 0))

    def set_ycoarse(self, on):
        """Per-radius y pass: hierarchical scan that skips row groups by their minima (default) / plain scan."""
        check(self.lib.psb200_set_option(self.handle, b"ycoarse", 1 if on else 0))

    def set_profile(self, on):
        check(self.lib.psb200_set_option(self.handle, b"profile", 1 if on else 0))

    def profile_read(self):
        """{kernel family: (total ms, launches)} since the last read (device-synchronising)."""
        n = self.lib.psb200_profile_kernels()
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int64 * n)()
        check(self.lib.psb200_profile_read(self.handle, ms, cnt))
        return {self.lib.psb200_profile_name(i).decode(): (ms[i], int(cnt[i])) for i in range(n) if cnt[i]}

    def profile_records(self, max_records=4096):
        """[(kernel family, ms)] per launch, in launch order (call before profile_read)."""
        ids = (ctypes.c_int * max_records)()
        ms = (ctypes.c_float * max_records)()
        n = min(self.lib.psb200_profile_records(self.handle, ids, ms, max_records), max_records)
        return [(self.lib.psb200_profile_name(ids[i]).decode(), float(ms[i])) for i in range(n)]

    def launch_count(self):
        return int(self.lib.psb200_launch_count(self.handle))

    def workspace(self, nbytes):
        """Grow-only device scratch buffer (a torch uint8 tensor used purely as an allocation)."""
        import torch
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=f"cuda:{self.device}")
        return self._ws

    def release_workspace(self):
        self._ws = None

    def __del__(self):
        try:
            if self.handle:
                self.lib.psb200_destroy(self.handle)
                self.handle = _vp()
        except Exception:
            pass


_contexts = {}


def context(device=None):
    import torch
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    device = int(device)
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]
