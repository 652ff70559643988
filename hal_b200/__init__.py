"""hal_b200 -- B200-native HAL liftover hot path.

The product is the C-ABI shared library ``hal_b200/libhalgpu.so`` (declared in ``include/halgpu.h``,
built for sm_100a by ``hal_b200/build.py``) plus the C++ host layer / CLIs under ``hal_b200/csrc/host``.
This module is only the ctypes binding the tests and ``bench.py`` drive it through; there is no CPU
implementation behind it -- without the CUDA library, or without a GPU, every call raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhalgpu.so")

HALGPU_NO_DUPES = 1
HALGPU_NO_SORT = 2
HALGPU_PSL = 4
HALGPU_COLUMN_LIFTOVER = 8
HALGPU_RAW_FRAGMENTS = 16
HALGPU_SEED_BOTTOM = 32
HALGPU_NO_FAST = 64
HALGPU_COUNT_DUPES = 1
HALGPU_NO_ANCESTORS = 2
HALGPU_COL_NO_DUPES = 4


class HalGpuError(RuntimeError):
    pass


class _Seq(C.Structure):
    _fields_ = [("name", C.c_char_p), ("start", C.c_int64), ("length", C.c_int64), ("num_top", C.c_int64),
                ("num_bottom", C.c_int64)]


class _Result(C.Structure):
    _fields_ = [("n", C.c_size_t), ("n_rec", C.c_size_t), ("offsets", C.c_void_p), ("recs", C.c_void_p),
                ("on_device", C.c_int), ("kernel_ms", C.c_float), ("launches", C.c_int), ("n_retry", C.c_size_t),
                ("psl", C.c_void_p), ("fast_ms", C.c_float), ("n_complex", C.c_size_t), ("owner", C.c_void_p)]


class _WigResult(C.Structure):
    _fields_ = [("n", C.c_size_t), ("pos", C.c_void_p), ("val", C.c_void_p), ("kernel_ms", C.c_float), ("launches", C.c_int),
                ("n_retry", C.c_size_t)]


REC_DTYPE = np.dtype([("start", "<i8"), ("end", "<i8"), ("src_start", "<i8"), ("tgt_seq", "<i4"),
                      ("strand", "u1"), ("src_strand", "u1"), ("n_frag", "<u2")])
assert REC_DTYPE.itemsize == 32
# HALGPU_RAW_FRAGMENTS: the same 32-byte slots hold halgpu_frag records (recs.view(FRAG_DTYPE))
FRAG_DTYPE = np.dtype([("src_start", "<i8"), ("tgt_start", "<i8"), ("length", "<i8"), ("flags", "<u4"), ("pad", "<u4")])
assert FRAG_DTYPE.itemsize == 32

# every symbol include/halgpu.h declares
ABI_SYMBOLS = [
    "halgpu_open", "halgpu_close", "halgpu_num_genomes", "halgpu_genome_name", "halgpu_genome_id",
    "halgpu_genome_parent", "halgpu_genome_num_children", "halgpu_genome_child", "halgpu_genome_length",
    "halgpu_genome_num_top", "halgpu_genome_num_bottom", "halgpu_newick", "halgpu_sequence_table", "halgpu_mrca",
    "halgpu_staged_bytes", "halgpu_stream", "halgpu_liftover", "halgpu_liftover_device", "halgpu_free_result",
    "halgpu_free_string", "halgpu_launch_count", "halgpu_columns_depth", "halgpu_columns_depth_device",
    "halgpu_column_runs", "halgpu_free_col_runs", "halgpu_genome_dna", "halgpu_host_alloc", "halgpu_host_free",
    "halgpu_comm_unique_id", "halgpu_comm_init", "halgpu_comm_free", "halgpu_comm_rank", "halgpu_comm_size",
    "halgpu_liftover_allgather_begin", "halgpu_liftover_allgather_end",
    "halgpu_maf_text", "halgpu_wiggle_liftover", "halgpu_free_wig_result", "halgpu_column_runs_in_sweep", "halgpu_genome_metadata", "halgpu_genome_top_segments", "halgpu_genome_bottom_segments",
]


def load_library(path=None):
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise HalGpuError(f"{path} is missing: build it with `python -m hal_b200.build` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.halgpu_open.argtypes = [C.c_char_p, i32, C.POINTER(vp), C.POINTER(C.c_char_p)]
    L.halgpu_close.argtypes = [vp]
    L.halgpu_num_genomes.argtypes = [vp]
    L.halgpu_genome_name.argtypes = [vp, i32]
    L.halgpu_genome_name.restype = C.c_char_p
    L.halgpu_genome_id.argtypes = [vp, C.c_char_p]
    for f in ("halgpu_genome_parent", "halgpu_genome_num_children"):
        getattr(L, f).argtypes = [vp, i32]
    L.halgpu_genome_child.argtypes = [vp, i32, i32]
    for f in ("halgpu_genome_length", "halgpu_genome_num_top", "halgpu_genome_num_bottom"):
        getattr(L, f).argtypes = [vp, i32]
        getattr(L, f).restype = i64
    L.halgpu_newick.argtypes = [vp]
    L.halgpu_newick.restype = C.c_char_p
    L.halgpu_sequence_table.argtypes = [vp, i32, C.POINTER(C.POINTER(_Seq)), C.POINTER(C.c_size_t)]
    L.halgpu_mrca.argtypes = [vp, i32, i32]
    L.halgpu_staged_bytes.argtypes = [vp]
    L.halgpu_staged_bytes.restype = C.c_size_t
    L.halgpu_stream.argtypes = [vp]
    L.halgpu_stream.restype = vp
    for f in ("halgpu_liftover", "halgpu_liftover_device"):
        getattr(L, f).argtypes = [vp, i32, i32, i32, C.c_uint32, C.c_size_t, vp, vp, vp,
                                  C.POINTER(C.POINTER(_Result)), C.POINTER(C.c_char_p)]
    for f in ("halgpu_columns_depth", "halgpu_columns_depth_device"):
        getattr(L, f).argtypes = [vp, i32, i64, i64, i64, vp, C.c_size_t, C.c_uint32, vp, C.POINTER(C.c_float),
                                  C.POINTER(C.c_char_p)]
    L.halgpu_wiggle_liftover.argtypes = [vp, i32, i32, C.c_uint32, C.c_size_t, vp, vp, vp, vp, C.c_size_t, C.c_size_t, vp, vp,
                                         C.POINTER(C.POINTER(_WigResult)), C.POINTER(C.c_char_p)]
    L.halgpu_free_wig_result.argtypes = [C.POINTER(_WigResult)]
    L.halgpu_host_alloc.argtypes = [C.c_size_t]
    L.halgpu_host_alloc.restype = vp
    L.halgpu_host_free.argtypes = [vp]
    L.halgpu_free_result.argtypes = [C.POINTER(_Result)]
    L.halgpu_free_string.argtypes = [C.c_void_p]
    L.halgpu_launch_count.restype = C.c_uint64
    L.halgpu_comm_unique_id.argtypes = [vp, C.POINTER(C.c_char_p)]
    L.halgpu_comm_init.argtypes = [vp, i32, i32, vp, C.POINTER(vp), C.POINTER(C.c_char_p)]
    L.halgpu_comm_free.argtypes = [vp]
    L.halgpu_comm_rank.argtypes = [vp]
    L.halgpu_comm_size.argtypes = [vp]
    L.halgpu_liftover_allgather_begin.argtypes = [vp, i32, i32, i32, C.c_uint32, C.c_size_t, vp, vp, vp, C.POINTER(vp), C.POINTER(C.c_char_p)]
    L.halgpu_liftover_allgather_end.argtypes = [vp, C.POINTER(C.POINTER(_Result)), vp, vp, C.POINTER(C.c_char_p)]
    return L


def _take_err(L, err):
    msg = err.value.decode() if err.value else "unknown error"
    return msg


class DeviceResult:
    """Result of liftover_device: device pointers, freed on close()."""

    def __init__(self, lib, res):
        self._lib, self._res = lib, res
        r = res.contents
        self.n, self.n_rec = r.n, r.n_rec
        self.offsets_ptr, self.recs_ptr = r.offsets, r.recs
        self.kernel_ms, self.launches, self.n_retry = r.kernel_ms, r.launches, r.n_retry
        self.fast_ms, self.n_complex = r.fast_ms, r.n_complex

    def close(self):
        if self._res is not None:
            self._lib.halgpu_free_result(self._res)
            self._res = None


class Alignment:
    """Read-only alignment staged in the HBM of one GPU (mirrors hal::Alignment's query surface for the path)."""

    def __init__(self, path, device=0, lib_path=None):
        self.L = load_library(lib_path)
        h, err = C.c_void_p(), C.c_char_p()
        # char** out-param: ctypes cannot free a c_char_p for us, keep the raw pointer
        errp = C.c_void_p()
        rc = self.L.halgpu_open(path.encode(), device, C.byref(h), C.cast(C.byref(errp), C.POINTER(C.c_char_p)))
        if rc != 0:
            msg = C.cast(errp, C.c_char_p).value.decode() if errp.value else "halgpu_open failed"
            if errp.value:
                self.L.halgpu_free_string(errp)
            raise HalGpuError(msg)
        self.h = h
        self.genomes = [self.L.halgpu_genome_name(h, g).decode() for g in range(self.L.halgpu_num_genomes(h))]

    def close(self):
        if getattr(self, "h", None):
            self.L.halgpu_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def genome_id(self, name):
        return self.L.halgpu_genome_id(self.h, name.encode())

    def genome_length(self, g):
        return self.L.halgpu_genome_length(self.h, g)

    def sequences(self, g):
        p, n = C.POINTER(_Seq)(), C.c_size_t()
        if self.L.halgpu_sequence_table(self.h, g, C.byref(p), C.byref(n)) != 0:
            raise HalGpuError("bad genome index")
        return [(p[i].name.decode(), p[i].start, p[i].length) for i in range(n.value)]

    @property
    def newick(self):
        return self.L.halgpu_newick(self.h).decode()

    @property
    def staged_bytes(self):
        return self.L.halgpu_staged_bytes(self.h)

    @property
    def stream(self):
        return self.L.halgpu_stream(self.h)

    def _call(self, fn, src, tgt, flags, n, a, b, s, coal=-1):
        res, errp = C.POINTER(_Result)(), C.c_void_p()
        rc = fn(self.h, src, tgt, coal, flags, n, a, b, s, C.byref(res), C.cast(C.byref(errp), C.POINTER(C.c_char_p)))
        if rc != 0:
            msg = C.cast(errp, C.c_char_p).value.decode() if errp.value else "liftover failed"
            if errp.value:
                self.L.halgpu_free_string(errp)
            raise HalGpuError(msg)
        return res

    def liftover(self, src, tgt, start, end_incl, strand=None, flags=0, coalescence_limit=-1):
        """Host arrays in (forward genome coordinates), numpy arrays out: (offsets[n+1], recs[REC_DTYPE], info).
        coalescence_limit: genome index for halLiftover --coalescenceLimit (-1: the MRCA)."""
        start = np.ascontiguousarray(start, dtype=np.int64)
        end_incl = np.ascontiguousarray(end_incl, dtype=np.int64)
        st = None if strand is None else np.ascontiguousarray(strand, dtype=np.uint8)
        n = len(start)
        res = self._call(self.L.halgpu_liftover, src, tgt, flags, n, start.ctypes.data, end_incl.ctypes.data,
                         None if st is None else st.ctypes.data, coalescence_limit)
        r = res.contents
        offsets = np.ctypeslib.as_array(C.cast(r.offsets, C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
        if r.n_rec:
            buf = (C.c_char * (r.n_rec * 32)).from_address(r.recs)
            recs = np.frombuffer(buf, dtype=REC_DTYPE).copy()
        else:
            recs = np.zeros(0, dtype=REC_DTYPE)
        info = dict(kernel_ms=r.kernel_ms, launches=r.launches, n_retry=r.n_retry, fast_ms=r.fast_ms, n_complex=r.n_complex)
        if r.psl and r.n_rec:
            info["psl"] = np.ctypeslib.as_array(C.cast(r.psl, C.POINTER(C.c_uint32)), shape=(r.n_rec, 4)).copy()
        self.L.halgpu_free_result(res)
        return offsets, recs, info

    def depth(self, ref, first, last, step=1, targets=(), flags=0, out_ptr=None):
        """halAlignmentDepth values for genome positions first..last (inclusive).  Returns (int32 array, kernel_ms);
        with out_ptr (a device pointer) the values stay on the GPU and None is returned in their place."""
        n = (last - first) // step + 1
        t = np.ascontiguousarray(list(targets), dtype=np.int32)
        ms, errp = C.c_float(0), C.c_void_p()
        out = None if out_ptr is not None else np.zeros(n, np.int32)
        fn = self.L.halgpu_columns_depth_device if out_ptr is not None else self.L.halgpu_columns_depth
        rc = fn(self.h, ref, first, last, step, t.ctypes.data if len(t) else None, len(t), flags,
                out_ptr if out_ptr is not None else out.ctypes.data, C.byref(ms), C.cast(C.byref(errp), C.POINTER(C.c_char_p)))
        if rc != 0:
            msg = C.cast(errp, C.c_char_p).value.decode() if errp.value else "depth failed"
            if errp.value:
                self.L.halgpu_free_string(errp)
            raise HalGpuError(msg)
        return out, ms.value

    def wiggle_liftover(self, src, tgt, run_first, run_last_incl, val_offset, vals, flags=0, preload_pos=(), preload_val=()):
        """halgpu_wiggle_liftover: runs of source bases with values in, (positions, values, info) of the target bases out."""
        f = np.ascontiguousarray(run_first, dtype=np.int64)
        l = np.ascontiguousarray(run_last_incl, dtype=np.int64)
        o = np.ascontiguousarray(val_offset, dtype=np.int64)
        v = np.ascontiguousarray(vals, dtype=np.float64)
        pp = np.ascontiguousarray(preload_pos, dtype=np.int64)
        pv = np.ascontiguousarray(preload_val, dtype=np.float64)
        res, errp = C.POINTER(_WigResult)(), C.c_void_p()
        rc = self.L.halgpu_wiggle_liftover(self.h, src, tgt, flags, len(f), f.ctypes.data, l.ctypes.data, o.ctypes.data, v.ctypes.data,
                                           len(v), len(pp), pp.ctypes.data if len(pp) else None, pv.ctypes.data if len(pp) else None,
                                           C.byref(res), C.cast(C.byref(errp), C.POINTER(C.c_char_p)))
        if rc != 0:
            msg = C.cast(errp, C.c_char_p).value.decode() if errp.value else "wiggle liftover failed"
            if errp.value:
                self.L.halgpu_free_string(errp)
            raise HalGpuError(msg)
        r = res.contents
        n = r.n
        pos = np.ctypeslib.as_array(C.cast(r.pos, C.POINTER(C.c_int64)), shape=(max(n, 1),))[:n].copy()
        val = np.ctypeslib.as_array(C.cast(r.val, C.POINTER(C.c_double)), shape=(max(n, 1),))[:n].copy()
        info = dict(kernel_ms=r.kernel_ms, launches=r.launches, n_retry=r.n_retry)
        self.L.halgpu_free_wig_result(res)
        return pos, val, info

    def liftover_ptrs(self, src, tgt, n, start_ptr, end_ptr, strand_ptr=None, flags=0, device=False):
        """Raw-pointer variants (pinned host buffers, or device buffers with device=True)."""
        fn = self.L.halgpu_liftover_device if device else self.L.halgpu_liftover
        res = self._call(fn, src, tgt, flags, n, start_ptr, end_ptr, strand_ptr)
        return DeviceResult(self.L, res)


class Comm:
    """One rank of a multi-GPU communicator over an Alignment (include/halgpu.h: halgpu_comm_*).  The 128-byte id comes
    from Comm.unique_id() on rank 0 and reaches the other ranks out of band (tests: shared between threads; bench.py:
    torch.distributed broadcast)."""

    def __init__(self, alignment, nranks, rank, uid):
        self.a, self.L = alignment, alignment.L
        h, errp = C.c_void_p(), C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(uid))
        rc = self.L.halgpu_comm_init(alignment.h, nranks, rank, C.addressof(buf), C.byref(h), C.cast(C.byref(errp), C.POINTER(C.c_char_p)))
        if rc != 0:
            raise HalGpuError(_err_text(self.L, errp, "halgpu_comm_init failed"))
        self.h, self.nranks, self.rank = h, nranks, rank

    @staticmethod
    def unique_id(lib):
        buf, errp = (C.c_uint8 * 128)(), C.c_void_p()
        if lib.halgpu_comm_unique_id(C.addressof(buf), C.cast(C.byref(errp), C.POINTER(C.c_char_p))) != 0:
            raise HalGpuError(_err_text(lib, errp, "halgpu_comm_unique_id failed"))
        return bytes(buf)

    def begin(self, src, tgt, n, start_ptr, end_ptr, strand_ptr=None, flags=0, coalescence_limit=-1):
        g, errp = C.c_void_p(), C.c_void_p()
        rc = self.L.halgpu_liftover_allgather_begin(self.h, src, tgt, coalescence_limit, flags, n, start_ptr, end_ptr, strand_ptr, C.byref(g),
                                                    C.cast(C.byref(errp), C.POINTER(C.c_char_p)))
        if rc != 0:
            raise HalGpuError(_err_text(self.L, errp, "allgather_begin failed"))
        return g

    def end(self, g):
        """-> (DeviceResult of the whole batch, intervals per rank, records per rank)"""
        res, errp = C.POINTER(_Result)(), C.c_void_p()
        npr, nrr = (C.c_size_t * self.nranks)(), (C.c_size_t * self.nranks)()
        rc = self.L.halgpu_liftover_allgather_end(g, C.byref(res), C.addressof(npr), C.addressof(nrr), C.cast(C.byref(errp), C.POINTER(C.c_char_p)))
        if rc != 0:
            raise HalGpuError(_err_text(self.L, errp, "allgather_end failed"))
        return DeviceResult(self.L, res), list(npr), list(nrr)

    def close(self):
        if getattr(self, "h", None):
            self.L.halgpu_comm_free(self.h)
            self.h = None


def _err_text(lib, errp, default):
    msg = C.cast(errp, C.c_char_p).value.decode() if errp.value else default
    if errp.value:
        lib.halgpu_free_string(errp)
    return msg
