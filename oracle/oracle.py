"""ctypes binding of the CPU ORACLE (oracle/pmc_oracle.c).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
SENTINEL = np.float32(1.0e18)


class Geom(C.Structure):
    _fields_ = [
        ("n_particles", C.c_int64), ("cps", C.c_int), ("n_cells", C.c_int64),
        ("nmax", C.c_int), ("n_M", C.c_int), ("w", C.c_float), ("L", C.c_float),
        ("half_L", C.c_float), ("sigma", C.c_float), ("sigma2", C.c_float),
        ("delta", C.c_float), ("dscale", C.c_float), ("L_box", C.c_double),
        ("seed", C.c_uint64), ("K", C.c_int), ("M", C.c_int), ("proposal", C.c_int), ("A", C.c_int),
    ]


class TrialRecord(C.Structure):
    _fields_ = [("sweep", C.c_uint64), ("cx", C.c_int32), ("cy", C.c_int32), ("slot", C.c_int32),
                ("cnt", C.c_int32), ("verdict", C.c_int32), ("trial", C.c_int32),
                ("px", C.c_float), ("py", C.c_float), ("own_x", C.c_float * 8), ("own_y", C.c_float * 8)]


def build(force=False):
    src = os.path.join(_HERE, "pmc_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        fp, sp = C.POINTER(C.c_float), C.POINTER(C.c_int16)
        gp, u64p = C.POINTER(Geom), C.POINTER(C.c_uint64)
        _lib.oracle_make_geom.argtypes = [C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int,
                                          C.c_int, C.c_float, C.c_uint64, C.c_int, gp]
        _lib.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        _lib.oracle_init_r.argtypes = [gp, fp]
        _lib.oracle_assign.argtypes = [gp, fp, fp, sp]
        _lib.oracle_assign.restype = C.c_int64
        _lib.oracle_cell_of.argtypes = [gp, C.c_float]
        _lib.oracle_subsweep.argtypes = [gp, fp, sp, C.POINTER(C.c_int), C.c_uint64, u64p, u64p]
        _lib.oracle_trial.argtypes = [gp, fp, sp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, fp]
        _lib.oracle_set_trace.argtypes = [C.c_void_p, C.c_int64]
        _lib.oracle_set_trace.restype = None
        _lib.oracle_trace_count.restype = C.c_int64
        _lib.oracle_shift_cells.argtypes = [gp, fp, sp, C.c_int, C.c_float]
        _lib.oracle_shift_cells.restype = C.c_int64
        _lib.oracle_schedule.argtypes = [gp, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int), fp]
        _lib.oracle_colour_to_off.argtypes = [C.c_int, C.POINTER(C.c_int)]
        _lib.oracle_sweep.argtypes = [gp, fp, sp, C.c_uint64, C.c_int, u64p, u64p]
        _lib.oracle_sweep.restype = C.c_int64
        _lib.oracle_sweep_omp.argtypes = [gp, fp, sp, C.c_uint64, C.c_int, u64p, u64p,
                                          C.POINTER(C.c_int64)]
        _lib.oracle_set_threads.argtypes = [C.c_int]
        _lib.oracle_disk_to_r.argtypes = [gp, fp, sp, fp]
        _lib.oracle_disk_to_r.restype = C.c_int64
        _lib.oracle_check.argtypes = [gp, fp, sp, C.POINTER(C.c_int64), fp]
        _lib.oracle_gr_hist.argtypes = [gp, fp, sp, C.c_float, C.c_int, u64p]
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _s(a):
    return a.ctypes.data_as(C.POINTER(C.c_int16))


def set_threads(n=0):
    """OpenMP threads of the timed CPU baseline; n = 0: every core this process may run on
    (sched_getaffinity), whatever OMP_NUM_THREADS says (torchrun exports 1)."""
    if n <= 0:
        n = len(os.sched_getaffinity(0))
    return int(lib().oracle_set_threads(n))


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return [int(x) for x in o]


class Oracle:
    """Host-side mirror of the C-ABI handle (include/pmc.h) on top of the C oracle."""

    def __init__(self, n_particles, phi=0.70, sigma_d=1.0, cell_w=2.0, nmax=8, n_M=4,
                 move_delta=0.1, seed=1234, cps_multiple=2, proposal=0):
        self.g = Geom()
        rc = lib().oracle_make_geom(n_particles, phi, sigma_d, cell_w, nmax, n_M, move_delta,
                                    seed, cps_multiple, C.byref(self.g))
        if rc:
            raise ValueError(f"oracle_make_geom failed rc={rc}")
        self.g.proposal = int(proposal)         # 0 uniform square (default), 1 the reference's Gaussian
        self.trials = C.c_uint64(0)
        self.accepted = C.c_uint64(0)
        self.lost = 0

    # geometry ---------------------------------------------------------------
    @property
    def n_cells(self):
        return int(self.g.n_cells)

    @property
    def cps(self):
        return int(self.g.cps)

    def alloc(self):
        disk = np.zeros((self.n_cells, 2, self.g.nmax), dtype=np.float32)
        n = np.zeros(self.n_cells, dtype=np.int16)
        return disk, n

    # path -------------------------------------------------------------------
    def init_r(self):
        r = np.zeros((2, self.g.n_particles), dtype=np.float32)
        if lib().oracle_init_r(C.byref(self.g), _f(r)):
            raise ValueError("n_particles is not a perfect square")
        return r

    def assign(self, r):
        r = np.ascontiguousarray(r, dtype=np.float32)
        disk, n = self.alloc()
        self.lost += lib().oracle_assign(C.byref(self.g), _f(r), _f(disk), _s(n))
        return disk, n

    def cell_of(self, x):
        return lib().oracle_cell_of(C.byref(self.g), C.c_float(x))

    def subsweep(self, disk, n, off, sweep):
        o = (C.c_int * 2)(*off)
        lib().oracle_subsweep(C.byref(self.g), _f(disk), _s(n), o, sweep,
                              C.byref(self.trials), C.byref(self.accepted))

    def trial(self, disk, n, cx, cy, slot, px, py):
        """(verdict, min_d2) of one trial: 0 accepted, 1 out of bound, 2 overlap (oracle_trial)."""
        md2 = C.c_float(np.float32(3.4e38))
        v = lib().oracle_trial(C.byref(self.g), _f(disk), _s(n), cx, cy, slot, C.c_float(px), C.c_float(py),
                               C.byref(md2))
        return v, np.float32(md2.value)

    def trace_on(self, cap):
        self._trace = (TrialRecord * cap)()
        lib().oracle_set_trace(C.cast(self._trace, C.c_void_p), cap)

    def trace_off(self):
        k = min(int(lib().oracle_trace_count()), len(self._trace))
        lib().oracle_set_trace(None, 0)
        return [self._trace[i] for i in range(k)]

    def shift_cells(self, disk, n, f, d):
        self.lost += lib().oracle_shift_cells(C.byref(self.g), _f(disk), _s(n), f, C.c_float(d))

    def schedule(self, sweep):
        order = (C.c_int * 4)()
        f = C.c_int()
        d = C.c_float()
        lib().oracle_schedule(C.byref(self.g), sweep, order, C.byref(f), C.byref(d))
        return list(order), f.value, np.float32(d.value)

    @staticmethod
    def colour_to_off(colour):
        o = (C.c_int * 2)()
        lib().oracle_colour_to_off(colour, o)
        return [o[0], o[1]]

    def sweep(self, disk, n, sweep0, n_sweeps, omp=False):
        if omp:
            lost = C.c_int64(0)
            threads = lib().oracle_sweep_omp(C.byref(self.g), _f(disk), _s(n), sweep0, n_sweeps,
                                             C.byref(self.trials), C.byref(self.accepted),
                                             C.byref(lost))
            self.lost += lost.value
            return threads
        self.lost += lib().oracle_sweep(C.byref(self.g), _f(disk), _s(n), sweep0, n_sweeps,
                                        C.byref(self.trials), C.byref(self.accepted))
        return 1

    def disk_to_r(self, disk, n):
        r = np.zeros((2, self.g.n_particles), dtype=np.float32)
        k = lib().oracle_disk_to_r(C.byref(self.g), _f(disk), _s(n), _f(r))
        return r, int(k)

    def check(self, disk, n):
        out = (C.c_int64 * 4)()
        md2 = C.c_float()
        lib().oracle_check(C.byref(self.g), _f(disk), _s(n), out, C.byref(md2))
        return {"total": out[0], "out_of_cell": out[1], "overlaps": out[2],
                "bad_sentinels": out[3], "min_d2": np.float32(md2.value)}

    def gr_hist(self, disk, n, r_max, nbins):
        h = np.zeros(nbins, dtype=np.uint64)
        lib().oracle_gr_hist(C.byref(self.g), _f(disk), _s(n), C.c_float(r_max), nbins,
                             h.ctypes.data_as(C.POINTER(C.c_uint64)))
        return h


# ---------------------------------------------------------------------------------------------------------
# 3-D Lennard-Jones mode (oracle/pmc_oracle_lj.c)
class LjGeom(C.Structure):
    _fields_ = [("n_particles", C.c_int64), ("n_cells", C.c_int64), ("cps", C.c_int), ("nmax", C.c_int), ("n_M", C.c_int),
                ("proposal", C.c_int), ("L", C.c_float), ("half_L", C.c_float), ("w", C.c_float), ("rc2", C.c_float),
                ("beta", C.c_float), ("sigma", C.c_float), ("dscale", C.c_float), ("seed", C.c_uint64)]


def _lj():
    L = lib()
    if getattr(L, "_lj_ready", False):
        return L
    gp, fp, sp, u64p = C.POINTER(LjGeom), C.POINTER(C.c_float), C.POINTER(C.c_int16), C.POINTER(C.c_uint64)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.oracle_lj_make_geom.argtypes = [C.c_int64, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint64, C.c_int, gp]
    L.oracle_lj_init_r.argtypes = [gp, fp]
    L.oracle_lj_assign.argtypes = [gp, fp, fp, sp]
    L.oracle_lj_assign.restype = C.c_int64
    L.pmc_lj_pair.argtypes = [C.c_float] * 4
    L.pmc_lj_pair.restype = C.c_float
    L.pmc_exp_det.argtypes = [C.c_float]
    L.pmc_exp_det.restype = C.c_float
    L.oracle_lj_subsweep.argtypes = [gp, fp, sp, ip, C.c_uint64, u64p, u64p, dp]
    L.oracle_lj_shift_cells.argtypes = [gp, fp, sp, C.c_int, C.c_float]
    L.oracle_lj_shift_cells.restype = C.c_int64
    L.oracle_lj_schedule.argtypes = [gp, C.c_uint64, ip, ip, fp]
    L.oracle_lj_colour_to_off.argtypes = [C.c_int, ip]
    L.oracle_lj_sweep.argtypes = [gp, fp, sp, C.c_uint64, C.c_int, u64p, u64p, dp]
    L.oracle_lj_sweep.restype = C.c_int64
    L.oracle_lj_energy.argtypes = [gp, fp, sp]
    L.oracle_lj_energy.restype = C.c_double
    L.oracle_lj_probe.argtypes = [gp, fp, sp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, ip, fp, fp]
    L._lj_ready = True
    return L


class OracleLJ:
    """Host-side mirror of include/pmc_lj.h on top of the C oracle."""

    def __init__(self, n_particles, L=10.0, beta=0.3, cells_per_side=4, nmax=10, n_M=10, sigma=0.5, seed=1234, proposal=0):
        self.g = LjGeom()
        if _lj().oracle_lj_make_geom(n_particles, L, beta, cells_per_side, nmax, n_M, sigma, seed, proposal, C.byref(self.g)):
            raise ValueError("oracle_lj_make_geom failed")
        self.trials, self.accepted, self.dE, self.lost = C.c_uint64(0), C.c_uint64(0), C.c_double(0.0), 0

    def init_r(self):
        r = np.zeros((3, self.g.n_particles), dtype=np.float32)
        if _lj().oracle_lj_init_r(C.byref(self.g), _f(r)):
            raise ValueError("n_particles is not a perfect cube")
        return r

    def assign(self, r):
        r = np.ascontiguousarray(r, dtype=np.float32)
        disk = np.zeros((self.g.n_cells, 3, self.g.nmax), dtype=np.float32)
        n = np.zeros(self.g.n_cells, dtype=np.int16)
        self.lost += _lj().oracle_lj_assign(C.byref(self.g), _f(r), _f(disk), _s(n))
        return disk, n

    def subsweep(self, disk, n, off, sweep):
        _lj().oracle_lj_subsweep(C.byref(self.g), _f(disk), _s(n), (C.c_int * 3)(*off), sweep, C.byref(self.trials),
                                 C.byref(self.accepted), C.byref(self.dE))

    def shift_cells(self, disk, n, f, d):
        self.lost += _lj().oracle_lj_shift_cells(C.byref(self.g), _f(disk), _s(n), f, C.c_float(d))

    def schedule(self, sweep):
        order, f, d = (C.c_int * 8)(), C.c_int(), C.c_float()
        _lj().oracle_lj_schedule(C.byref(self.g), sweep, order, C.byref(f), C.byref(d))
        return list(order), f.value, np.float32(d.value)

    @staticmethod
    def colour_to_off(colour):
        o = (C.c_int * 3)()
        _lj().oracle_lj_colour_to_off(colour, o)
        return [o[0], o[1], o[2]]

    def sweep(self, disk, n, sweep0, n_sweeps):
        tr = np.zeros(max(n_sweeps, 1), dtype=np.float64)
        self.lost += _lj().oracle_lj_sweep(C.byref(self.g), _f(disk), _s(n), sweep0, n_sweeps, C.byref(self.trials),
                                           C.byref(self.accepted), tr.ctypes.data_as(C.POINTER(C.c_double)))
        return tr[:n_sweeps]

    def energy(self, disk, n):
        return float(_lj().oracle_lj_energy(C.byref(self.g), _f(disk), _s(n)))

    def probe(self, disk, n, cx, cy, cz, slot, px, py, pz):
        oob, ec, en = C.c_int(), C.c_float(), C.c_float()
        _lj().oracle_lj_probe(C.byref(self.g), _f(disk), _s(n), cx, cy, cz, slot, C.c_float(px), C.c_float(py), C.c_float(pz),
                              C.byref(oob), C.byref(ec), C.byref(en))
        return oob.value, np.float32(ec.value), np.float32(en.value)
