"""parallel-monte-carlo_b200: host-side mirror of the reference driver's call sites.

The product is `libpmc_b200.so` (hand-written sm_100a CUDA behind the C-ABI of
include/pmc.h).  This module only binds it with ctypes and uses torch for device memory,
streams and torch.distributed plumbing.  There is NO CPU fallback: if the library is not
built or no GPU is present, the compute entry points raise.

Reference call sites mirrored (qingye3/parallel-monte-carlo, start.cu:main):
    init_r   start.cu:212      -> ParallelMC.init_r
    assign   start.cu:227      -> ParallelMC.assign
    subsweep start.cu:242-245  -> ParallelMC.subsweep
    shift    start.cu:255      -> ParallelMC.shift_cells
    loop     start.cu:237-260  -> ParallelMC.sweep
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("PMC_LIB_PATH") or os.path.join(_HERE, "libpmc_b200.so")     # the override is a development aid
SOURCES = ["pmc_api.cu", "pmc_cells.cu", "pmc_sweep.cu", "pmc_sweep4.cu", "pmc_lj.cu"]
HEADERS = ["pmc_internal.cuh", os.path.join("..", "..", "include", "pmc.h"), os.path.join("..", "..", "include", "pmc_lj.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

PMC_E = {-1: "INVALID", -2: "UNSUPPORTED", -3: "OVERFLOW", -4: "LOST", -5: "NOT_SQUARE", -6: "COMM"}
STATUS_OVERFLOW, STATUS_LOST = 1, 2


class PmcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pmc error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("n_particles", C.c_int64), ("phi", C.c_float), ("sigma_d", C.c_float),
                ("cell_w", C.c_float), ("nmax", C.c_int), ("n_M", C.c_int),
                ("move_delta", C.c_float), ("seed", C.c_uint64), ("cps_multiple", C.c_int),
                ("device", C.c_int), ("rank", C.c_int), ("n_ranks", C.c_int), ("proposal", C.c_int)]


class Geometry(C.Structure):
    _fields_ = [("n_particles", C.c_int64), ("cps", C.c_int), ("n_cells", C.c_int64),
                ("nmax", C.c_int), ("n_M", C.c_int), ("w", C.c_float), ("L", C.c_float),
                ("sigma_d", C.c_float), ("move_delta", C.c_float), ("row0", C.c_int),
                ("rows", C.c_int), ("ghost_rows", C.c_int), ("local_cells", C.c_int64), ("grid_q", C.c_float)]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_HERE, "csrc", s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile libpmc_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + ["-o", LIB_PATH] + [os.path.join(_HERE, "csrc", s) for s in SOURCES]
    cmd += ["-lcudart", "-ldl"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return LIB_PATH


_lib = None


def lib():
    """Load the C-ABI library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built: run __graft_entry__.build() (needs nvcc); "
                          "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    hp = C.c_void_p
    vp = C.c_void_p
    L.pmc_create.argtypes = [C.POINTER(Params), C.POINTER(hp)]
    L.pmc_destroy.argtypes = [hp]
    L.pmc_get_geometry.argtypes = [hp, C.POINTER(Geometry)]
    for f in ("pmc_r_bytes", "pmc_disk_bytes", "pmc_n_bytes"):
        getattr(L, f).argtypes = [hp]
        getattr(L, f).restype = C.c_size_t
    L.pmc_set_stream.argtypes = [hp, vp]
    L.pmc_set_blocking.argtypes = [hp, C.c_int]
    L.pmc_synchronize.argtypes = [hp]
    L.pmc_set_tuning.argtypes = [hp, C.c_char_p, C.c_int]
    L.pmc_error_string.argtypes = [C.c_int]
    L.pmc_error_string.restype = C.c_char_p
    L.pmc_init_r.argtypes = [hp, vp]
    L.pmc_assign.argtypes = [hp, vp, vp, vp]
    L.pmc_subsweep.argtypes = [hp, vp, vp, C.POINTER(C.c_int), C.c_uint64]
    L.pmc_shift_cells.argtypes = [hp, vp, vp, C.c_int, C.c_float]
    L.pmc_schedule.argtypes = [hp, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float)]
    L.pmc_colour_to_off.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.pmc_colour_to_off.restype = None
    L.pmc_plan_sweep.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_float, C.POINTER(C.c_int)]
    L.pmc_sweep.argtypes = [hp, vp, vp, C.c_uint64, C.c_int]
    L.pmc_get_counters.argtypes = [hp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                   C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    L.pmc_reset_counters.argtypes = [hp]
    L.pmc_get_kernel_time.argtypes = [hp, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    L.pmc_get_launch_count.argtypes = [hp, C.POINTER(C.c_longlong)]
    L.pmc_check.argtypes = [hp, vp, vp, C.POINTER(C.c_int64), C.POINTER(C.c_float)]
    L.pmc_gr_hist.argtypes = [hp, vp, vp, C.c_float, C.c_int, vp]
    L.pmc_pressure_from_hist.argtypes = [hp, vp, C.c_float, C.c_int, C.c_int64, vp,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.pmc_disk_to_r_host.argtypes = [hp, vp, vp, vp, C.POINTER(C.c_int64)]
    L.pmc_disk_to_r.argtypes = [hp, vp, vp, vp, C.POINTER(C.c_int64)]
    L.pmc_run_host.argtypes = [hp, vp, C.c_uint64, C.c_int, vp, vp]
    L.pmc_geometry_from_params.argtypes = [C.POINTER(Params), C.POINTER(Geometry)]
    L.pmc_rsa_host.argtypes = [C.POINTER(Params), C.c_uint64, vp, C.POINTER(C.c_int64)]
    L.pmc_write_dump.argtypes = [hp, vp, vp, C.c_char_p, C.c_int, C.c_int]
    L.pmc_save_checkpoint.argtypes = [hp, vp, vp, C.c_uint64, C.c_char_p]
    L.pmc_load_checkpoint.argtypes = [hp, C.c_char_p, vp, vp, C.POINTER(C.c_uint64)]
    L.pmc_comm_unique_id.argtypes = [vp]
    L.pmc_comm_init.argtypes = [hp, vp]
    L.pmc_exchange_ghosts.argtypes = [hp, vp, vp]
    _lib = L
    return L


EXPORTS = ["pmc_create", "pmc_destroy", "pmc_get_geometry", "pmc_r_bytes", "pmc_disk_bytes",
           "pmc_n_bytes", "pmc_set_stream", "pmc_set_blocking", "pmc_synchronize", "pmc_set_tuning",
           "pmc_error_string", "pmc_init_r", "pmc_assign", "pmc_subsweep", "pmc_shift_cells",
           "pmc_schedule", "pmc_colour_to_off", "pmc_plan_sweep", "pmc_sweep", "pmc_get_counters",
           "pmc_reset_counters", "pmc_get_kernel_time", "pmc_get_launch_count", "pmc_check", "pmc_gr_hist", "pmc_pressure_from_hist",
           "pmc_disk_to_r_host", "pmc_disk_to_r", "pmc_run_host", "pmc_geometry_from_params", "pmc_rsa_host",
           "pmc_write_dump", "pmc_save_checkpoint", "pmc_load_checkpoint", "pmc_comm_unique_id", "pmc_comm_init",
           "pmc_exchange_ghosts"]


def plan_sweep(order, f, d):
    """Tile extent, halo and per-colour margins the fused sweep uses for this colour order and shift
    (pmc_plan_sweep: host only)."""
    o = (C.c_int * 4)(*order)
    out = (C.c_int * 12)()
    rc = lib().pmc_plan_sweep(o, int(f), float(d), out)
    if rc:
        raise PmcError(rc, "pmc_plan_sweep")
    v = list(out)
    return dict(tx=v[0], ty=v[1], hx=v[2], hy=v[3], lo_x=v[4:8], lo_y=v[8:12])


def geometry_from_params(n_particles, phi=0.70, sigma_d=1.0, cell_w=2.0, nmax=8, n_M=4, move_delta=0.1,
                         seed=1234, cps_multiple=2, rank=0, n_ranks=1):
    """Geometry derivation without a device (pure host maths in the C library)."""
    p = Params(n_particles, phi, sigma_d, cell_w, nmax, n_M, move_delta, seed, cps_multiple, -1, rank, n_ranks, 0)
    g = Geometry()
    _ck(lib().pmc_geometry_from_params(C.byref(p), C.byref(g)))
    return g


def rsa_host(n_particles, seed=1, **kw):
    """Random sequential addition on the host (no device needed): numpy r [2][N], attempts."""
    import numpy as np
    p = Params(n_particles, kw.get("phi", 0.30), kw.get("sigma_d", 1.0), kw.get("cell_w", 2.0), 8, 4,
               kw.get("move_delta", 0.1), 1234, kw.get("cps_multiple", 2), -1, 0, 1, 0)
    r = np.zeros((2, n_particles), dtype=np.float32)
    att = C.c_int64()
    _ck(lib().pmc_rsa_host(C.byref(p), seed, r.ctypes.data, C.byref(att)))
    return r, att.value


def _ck(rc):
    if rc != 0:
        raise PmcError(rc, lib().pmc_error_string(rc).decode())


class ParallelMC:
    """One simulation handle on one GPU (or one slab of a multi-GPU run).

    strict (default True): a blocking call during which a cell overflowed nmax or particles fell outside the
    box raises PmcError (PMC_E_OVERFLOW / PMC_E_LOST).  With strict = False the code is kept in
    `last_warning` instead (the reference itself writes past nmax silently and carries on); the totals are in
    counters() either way."""
    strict = True
    last_warning = 0

    def _ckw(self, rc):
        if rc in (-3, -4) and not self.strict:
            self.last_warning = rc
            return
        _ck(rc)

    def __init__(self, n_particles, phi=0.70, sigma_d=1.0, cell_w=2.0, nmax=8, n_M=4,
                 move_delta=0.1, seed=1234, cps_multiple=2, device=-1, rank=0, n_ranks=1, proposal=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("parallel-monte-carlo_b200 needs a CUDA device (no CPU fallback)")
        self.torch = torch
        self.params = Params(n_particles, phi, sigma_d, cell_w, nmax, n_M, move_delta, seed,
                             cps_multiple, device, rank, n_ranks, proposal)
        self._h = C.c_void_p()
        _ck(lib().pmc_create(C.byref(self.params), C.byref(self._h)))
        self.geom = Geometry()
        _ck(lib().pmc_get_geometry(self._h, C.byref(self.geom)))
        self.device = torch.device("cuda", torch.cuda.current_device() if device < 0 else device)
        self.use_torch_stream()

    def close(self):
        if self._h:
            lib().pmc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # plumbing -----------------------------------------------------------------
    def use_torch_stream(self):
        s = self.torch.cuda.current_stream(self.device)
        _ck(lib().pmc_set_stream(self._h, C.c_void_p(s.cuda_stream)))

    def set_blocking(self, blocking):
        _ck(lib().pmc_set_blocking(self._h, int(blocking)))

    def synchronize(self):
        _ck(lib().pmc_synchronize(self._h))

    def set_tuning(self, name, value):
        """Result-invariant knobs (pmc_set_tuning): bands, prefetch, overlap, generic, four_plane, force_crowded, no_ns4,
        full_halo, tile_rows."""
        _ck(lib().pmc_set_tuning(self._h, name.encode(), int(value)))

    def alloc_r(self):
        return self.torch.empty((2, self.geom.n_particles), dtype=self.torch.float32, device=self.device)

    def alloc_cells(self):
        t = self.torch
        disk = t.empty((self.geom.local_cells, 2, self.geom.nmax), dtype=t.float32, device=self.device)
        n = t.empty((self.geom.local_cells,), dtype=t.int16, device=self.device)
        return disk, n

    # the four call sites ------------------------------------------------------
    def init_r(self, r=None):
        r = self.alloc_r() if r is None else r
        _ck(lib().pmc_init_r(self._h, r.data_ptr()))
        return r

    def assign(self, r, disk=None, n=None):
        if disk is None:
            disk, n = self.alloc_cells()
        self._ckw(lib().pmc_assign(self._h, r.data_ptr(), disk.data_ptr(), n.data_ptr()))
        return disk, n

    def subsweep(self, disk, n, off, sweep):
        o = (C.c_int * 2)(*off)
        self._ckw(lib().pmc_subsweep(self._h, disk.data_ptr(), n.data_ptr(), o, sweep))

    def shift_cells(self, disk, n, f, d):
        self._ckw(lib().pmc_shift_cells(self._h, disk.data_ptr(), n.data_ptr(), f, C.c_float(d)))

    def schedule(self, sweep):
        order = (C.c_int * 4)()
        f = C.c_int()
        d = C.c_float()
        _ck(lib().pmc_schedule(self._h, sweep, order, C.byref(f), C.byref(d)))
        return list(order), f.value, d.value

    @staticmethod
    def colour_to_off(colour):
        o = (C.c_int * 2)()
        lib().pmc_colour_to_off(colour, o)
        return [o[0], o[1]]

    def sweep(self, disk, n, sweep0, n_sweeps):
        self._ckw(lib().pmc_sweep(self._h, disk.data_ptr(), n.data_ptr(), sweep0, n_sweeps))

    # counters / observables ---------------------------------------------------
    def counters(self):
        tr, ac, lo = C.c_uint64(), C.c_uint64(), C.c_uint64()
        st = C.c_uint32()
        _ck(lib().pmc_get_counters(self._h, C.byref(tr), C.byref(ac), C.byref(lo), C.byref(st)))
        return {"trials": tr.value, "accepted": ac.value, "lost": lo.value, "status": st.value}

    def reset_counters(self):
        _ck(lib().pmc_reset_counters(self._h))

    def launch_count(self):
        nl = C.c_longlong()
        _ck(lib().pmc_get_launch_count(self._h, C.byref(nl)))
        return nl.value

    def kernel_time(self):
        """(ms, launches) of the fused sweep kernels since the last reset_counters()."""
        ms, nl = C.c_double(), C.c_longlong()
        _ck(lib().pmc_get_kernel_time(self._h, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def check(self, disk, n):
        out = (C.c_int64 * 4)()
        md2 = C.c_float()
        _ck(lib().pmc_check(self._h, disk.data_ptr(), n.data_ptr(), out, C.byref(md2)))
        return {"total": out[0], "out_of_cell": out[1], "overlaps": out[2],
                "bad_sentinels": out[3], "min_d2": md2.value}

    def gr_hist(self, disk, n, r_max, nbins):
        import numpy as np
        h = np.zeros(nbins, dtype=np.uint64)
        _ck(lib().pmc_gr_hist(self._h, disk.data_ptr(), n.data_ptr(), C.c_float(r_max), nbins,
                              h.ctypes.data))
        return h

    def pressure_from_hist(self, hist, r_max, n_samples=1):
        import numpy as np
        hist = np.ascontiguousarray(hist, dtype=np.uint64)
        g = np.zeros(len(hist), dtype=np.float64)
        gc, bp = C.c_double(), C.c_double()
        _ck(lib().pmc_pressure_from_hist(self._h, hist.ctypes.data, C.c_float(r_max), len(hist),
                                         n_samples, g.ctypes.data, C.byref(gc), C.byref(bp)))
        return g, gc.value, bp.value

    # host I/O -----------------------------------------------------------------
    def disk_to_r_host(self, disk, n):
        import numpy as np
        r = np.zeros((2, self.geom.n_particles), dtype=np.float32)
        k = C.c_int64()
        _ck(lib().pmc_disk_to_r_host(self._h, disk.data_ptr(), n.data_ptr(), r.ctypes.data, C.byref(k)))
        return r, k.value

    def disk_to_r(self, disk, n, r=None):
        """disk_to_r (kernel.cu:497-507) on the device: (r [2][N] global coordinates, particles found)."""
        r = self.alloc_r() if r is None else r
        k = C.c_int64()
        _ck(lib().pmc_disk_to_r(self._h, disk.data_ptr(), n.data_ptr(), r.data_ptr(), C.byref(k)))
        return r, k.value

    def run_host(self, r_host, sweep0, n_sweeps, disk_host, n_host):
        """End to end with host buffers (torch CPU tensors, ideally pinned)."""
        self._ckw(lib().pmc_run_host(self._h, r_host.data_ptr(), sweep0, n_sweeps,
                               disk_host.data_ptr(), n_host.data_ptr()))

    # initial configurations / trajectory / checkpoint ---------------------------
    def rsa(self, seed=1):
        """Random-sequential-addition configuration (host), uploaded as r [2][N]."""
        import numpy as np
        r = np.zeros((2, self.geom.n_particles), dtype=np.float32)
        att = C.c_int64()
        _ck(lib().pmc_rsa_host(C.byref(self.params), seed, r.ctypes.data, C.byref(att)))
        self.rsa_attempts = att.value
        return self.torch.from_numpy(r).to(self.device)

    def write_dump(self, disk, n, path, timestep=0, append=False):
        _ck(lib().pmc_write_dump(self._h, disk.data_ptr(), n.data_ptr(), path.encode(), timestep, int(append)))

    def save_checkpoint(self, disk, n, sweep, path):
        _ck(lib().pmc_save_checkpoint(self._h, disk.data_ptr(), n.data_ptr(), sweep, path.encode()))

    def load_checkpoint(self, path, disk=None, n=None):
        if disk is None:
            disk, n = self.alloc_cells()
        sw = C.c_uint64()
        _ck(lib().pmc_load_checkpoint(self._h, path.encode(), disk.data_ptr(), n.data_ptr(), C.byref(sw)))
        return disk, n, sw.value

    # multi-GPU ----------------------------------------------------------------
    def comm_init_from_torch(self):
        """Create the NCCL ring of this handle; the unique id travels over torch.distributed."""
        import torch.distributed as dist
        t = self.torch
        buf = (C.c_char * 128)()
        if dist.get_rank() == 0:
            _ck(lib().pmc_comm_unique_id(buf))
        idt = t.frombuffer(bytearray(bytes(buf)), dtype=t.uint8).clone()
        if dist.get_backend() == "nccl":
            idt = idt.to(self.device)
        dist.broadcast(idt, src=0)
        raw = bytes(idt.cpu().numpy().tobytes())
        cbuf = (C.c_char * 128).from_buffer_copy(raw)
        _ck(lib().pmc_comm_init(self._h, cbuf))

    def exchange_ghosts(self, disk, n):
        _ck(lib().pmc_exchange_ghosts(self._h, disk.data_ptr(), n.data_ptr()))


# ---------------------------------------------------------------------------------------------------------
# 3-D Lennard-Jones mode (include/pmc_lj.h): the reference's actual physics behind the same call sites
class LjParams(C.Structure):
    _fields_ = [("n_particles", C.c_int64), ("L", C.c_float), ("beta", C.c_float), ("cells_per_side", C.c_int),
                ("nmax", C.c_int), ("n_M", C.c_int), ("sigma", C.c_float), ("seed", C.c_uint64),
                ("proposal", C.c_int), ("device", C.c_int)]


LJ_EXPORTS = ["pmc_lj_create", "pmc_lj_destroy", "pmc_lj_r_bytes", "pmc_lj_disk_bytes", "pmc_lj_n_bytes", "pmc_lj_set_stream",
              "pmc_lj_init_r", "pmc_lj_assign", "pmc_lj_subsweep", "pmc_lj_shift_cells", "pmc_lj_schedule",
              "pmc_lj_colour_to_off", "pmc_lj_sweep", "pmc_lj_energy", "pmc_lj_get_counters", "pmc_lj_reset_counters",
              "pmc_lj_disk_to_r_host"]


def _lj_lib():
    L = lib()
    if getattr(L, "_lj_ready", False):
        return L
    hp, vp = C.c_void_p, C.c_void_p
    L.pmc_lj_create.argtypes = [C.POINTER(LjParams), C.POINTER(hp)]
    L.pmc_lj_destroy.argtypes = [hp]
    for f in ("pmc_lj_r_bytes", "pmc_lj_disk_bytes", "pmc_lj_n_bytes"):
        getattr(L, f).argtypes = [hp]
        getattr(L, f).restype = C.c_size_t
    L.pmc_lj_set_stream.argtypes = [hp, vp]
    L.pmc_lj_init_r.argtypes = [hp, vp]
    L.pmc_lj_assign.argtypes = [hp, vp, vp, vp]
    L.pmc_lj_subsweep.argtypes = [hp, vp, vp, C.POINTER(C.c_int), C.c_uint64]
    L.pmc_lj_shift_cells.argtypes = [hp, vp, vp, C.c_int, C.c_float]
    L.pmc_lj_schedule.argtypes = [hp, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float)]
    L.pmc_lj_colour_to_off.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.pmc_lj_colour_to_off.restype = None
    L.pmc_lj_sweep.argtypes = [hp, vp, vp, C.c_uint64, C.c_int, vp]
    L.pmc_lj_energy.argtypes = [hp, vp, vp, C.POINTER(C.c_double)]
    L.pmc_lj_get_counters.argtypes = [hp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
    L.pmc_lj_reset_counters.argtypes = [hp]
    L.pmc_lj_disk_to_r_host.argtypes = [hp, vp, vp, vp, C.POINTER(C.c_int64)]
    L._lj_ready = True
    return L


class ParallelMCLJ:
    """3-D Lennard-Jones handle: the reference's own arrays (r [3][N], disk [cells][3][nmax] global
    coordinates, int16 n) and call sites (start.cu:212,227,242-245,255), V2 energy accounting."""
    strict = True
    last_warning = 0

    def __init__(self, n_particles, L=10.0, beta=0.3, cells_per_side=4, nmax=10, n_M=10, sigma=0.5, seed=1234,
                 proposal=0, device=-1):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("parallel-monte-carlo_b200 needs a CUDA device (no CPU fallback)")
        self.torch = torch
        self.params = LjParams(n_particles, L, beta, cells_per_side, nmax, n_M, sigma, seed, proposal, device)
        self._h = C.c_void_p()
        _ck(_lj_lib().pmc_lj_create(C.byref(self.params), C.byref(self._h)))
        self.device = torch.device("cuda", torch.cuda.current_device() if device < 0 else device)
        self.n_cells = cells_per_side ** 3
        _ck(_lj_lib().pmc_lj_set_stream(self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def close(self):
        if self._h:
            _lj_lib().pmc_lj_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ckw(self, rc):
        if rc in (-3, -4) and not self.strict:
            self.last_warning = rc
            return
        _ck(rc)

    def alloc_cells(self):
        t = self.torch
        return (t.zeros((self.n_cells, 3, self.params.nmax), dtype=t.float32, device=self.device),
                t.zeros((self.n_cells,), dtype=t.int16, device=self.device))

    def init_r(self):
        r = self.torch.empty((3, self.params.n_particles), dtype=self.torch.float32, device=self.device)
        _ck(_lj_lib().pmc_lj_init_r(self._h, r.data_ptr()))
        return r

    def assign(self, r):
        disk, n = self.alloc_cells()
        self._ckw(_lj_lib().pmc_lj_assign(self._h, r.data_ptr(), disk.data_ptr(), n.data_ptr()))
        return disk, n

    def subsweep(self, disk, n, off, sweep):
        o = (C.c_int * 3)(*off)
        self._ckw(_lj_lib().pmc_lj_subsweep(self._h, disk.data_ptr(), n.data_ptr(), o, sweep))

    def shift_cells(self, disk, n, f, d):
        self._ckw(_lj_lib().pmc_lj_shift_cells(self._h, disk.data_ptr(), n.data_ptr(), f, C.c_float(d)))

    def schedule(self, sweep):
        order, f, d = (C.c_int * 8)(), C.c_int(), C.c_float()
        _ck(_lj_lib().pmc_lj_schedule(self._h, sweep, order, C.byref(f), C.byref(d)))
        return list(order), f.value, d.value

    @staticmethod
    def colour_to_off(colour):
        o = (C.c_int * 3)()
        _lj_lib().pmc_lj_colour_to_off(colour, o)
        return [o[0], o[1], o[2]]

    def sweep(self, disk, n, sweep0, n_sweeps, trace=False):
        """n_sweeps x (8 sub-sweeps + shiftCells); trace=True returns the accepted energy change of every sweep."""
        import numpy as np
        tr = np.zeros(max(n_sweeps, 1), dtype=np.float64) if trace else None
        self._ckw(_lj_lib().pmc_lj_sweep(self._h, disk.data_ptr(), n.data_ptr(), sweep0, n_sweeps,
                                         tr.ctypes.data if trace else None))
        return tr[:n_sweeps] if trace else None

    def energy(self, disk, n):
        e = C.c_double()
        _ck(_lj_lib().pmc_lj_energy(self._h, disk.data_ptr(), n.data_ptr(), C.byref(e)))
        return e.value

    def counters(self):
        tr, ac, lo, st, de = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_double()
        _ck(_lj_lib().pmc_lj_get_counters(self._h, C.byref(tr), C.byref(ac), C.byref(lo), C.byref(st), C.byref(de)))
        return {"trials": tr.value, "accepted": ac.value, "lost": lo.value, "status": st.value, "dE": de.value}

    def reset_counters(self):
        _ck(_lj_lib().pmc_lj_reset_counters(self._h))

    def disk_to_r_host(self, disk, n):
        import numpy as np
        r = np.zeros((3, self.params.n_particles), dtype=np.float32)
        k = C.c_int64()
        _ck(_lj_lib().pmc_lj_disk_to_r_host(self._h, disk.data_ptr(), n.data_ptr(), r.ctypes.data, C.byref(k)))
        return r, k.value
