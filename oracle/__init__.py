"""ctypes loader for the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / ``--impl reference`` legs -- never by
``lancedb_b200`` (the product path).  PARITY STATUS: IVF_PQ parity unpinned
(see oracle/oracle.h); flat path pinned by the reference's doctests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

METRICS = {"l2": 0, "cosine": 1, "dot": 2}


class _Index(C.Structure):
    _fields_ = [
        ("dim", C.c_uint32), ("nlist", C.c_uint32), ("m", C.c_uint32), ("metric", C.c_int),
        ("nrows", C.c_uint64),
        ("centroids", C.c_void_p), ("codebook", C.c_void_p), ("part_offsets", C.c_void_p),
        ("codes_t", C.c_void_p), ("row_ids", C.c_void_p), ("vectors", C.c_void_p),
    ]


class _Params(C.Structure):
    _fields_ = [
        ("k", C.c_uint32), ("nprobes", C.c_uint32), ("refine_factor", C.c_uint32),
        ("has_lower", C.c_int), ("has_upper", C.c_int), ("lower", C.c_float), ("upper", C.c_float),
        ("allow", C.c_void_p), ("allow_bits", C.c_uint64), ("max_nprobes", C.c_uint32),
    ]


def allow_bitmap(row_ids, nbits: int) -> np.ndarray:
    """Row-id allow-list -> u32 bitmap of `nbits` bits (the form orc_params.allow / the C ABI take)."""
    bm = np.zeros((int(nbits) + 31) // 32, np.uint32)
    r = np.asarray(row_ids, np.uint64)
    r = r[r < nbits]
    np.bitwise_or.at(bm, (r >> np.uint64(5)).astype(np.int64), (np.uint32(1) << (r & np.uint64(31)).astype(np.uint32)))
    return bm


def _params(k, nprobes, refine_factor, lower, upper, allow=None, allow_bits=0, max_nprobes=0):
    p = _Params(k, nprobes, refine_factor, lower is not None, upper is not None,
                0.0 if lower is None else lower, 0.0 if upper is None else upper, None, 0, int(max_nprobes or 0))
    if allow is not None:
        p.allow = allow.ctypes.data
        p.allow_bits = int(allow_bits)
    return p


def build(force: bool = False) -> str:
    """Compile oracle.c with gcc (flags in oracle/Makefile)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "oracle.h"))
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(_LIB_PATH)
    f32p, vp = C.POINTER(C.c_float), C.c_void_p
    lib.orc_l2_f32.restype = C.c_float
    lib.orc_l2_f32.argtypes = [vp, vp, C.c_size_t]
    lib.orc_dot_f32.restype = C.c_float
    lib.orc_dot_f32.argtypes = [vp, vp, C.c_size_t]
    lib.orc_cosine_f32.restype = C.c_float
    lib.orc_cosine_f32.argtypes = [vp, vp, C.c_size_t]
    lib.orc_distance_f32.restype = C.c_float
    lib.orc_distance_f32.argtypes = [C.c_int, vp, vp, C.c_size_t]
    lib.orc_l2_subvec.restype = C.c_float
    lib.orc_l2_subvec.argtypes = [vp, vp, C.c_size_t]
    lib.orc_normalize_f32.restype = None
    lib.orc_normalize_f32.argtypes = [vp, C.c_size_t, vp]
    lib.orc_find_partitions.restype = None
    lib.orc_find_partitions.argtypes = [C.POINTER(_Index), vp, C.c_uint32, vp, vp, vp]
    lib.orc_build_lut.restype = None
    lib.orc_build_lut.argtypes = [C.POINTER(_Index), vp, vp]
    lib.orc_pq_scan.restype = None
    lib.orc_pq_scan.argtypes = [vp, vp, C.c_size_t, C.c_uint32, vp]
    lib.orc_partition_distances.restype = None
    lib.orc_partition_distances.argtypes = [C.POINTER(_Index), vp, C.c_uint32, vp]
    lib.orc_ivf_assign.argtypes = [C.POINTER(_Index), vp, C.c_uint64, vp]
    lib.orc_ivf_assign.restype = None
    lib.orc_pq_encode.argtypes = [C.POINTER(_Index), vp, vp, C.c_uint64, vp]
    lib.orc_pq_encode.restype = None
    lib.orc_ivfpq_search.restype = C.c_int
    lib.orc_ivfpq_search.argtypes = [C.POINTER(_Index), vp, C.c_uint32, C.POINTER(_Params), vp, vp, vp, C.c_int]
    lib.orc_flat_search.restype = C.c_int
    lib.orc_flat_search.argtypes = [vp, C.c_uint64, C.c_uint32, vp, C.c_int, vp, C.c_uint32,
                                    C.POINTER(_Params), vp, vp, vp, C.c_int]
    del f32p
    _lib = lib
    return lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def l2(x, y) -> float:
    x, y = _f32(x), _f32(y)
    return float(load().orc_l2_f32(_ptr(x), _ptr(y), x.size))


def dot(x, y) -> float:
    x, y = _f32(x), _f32(y)
    return float(load().orc_dot_f32(_ptr(x), _ptr(y), x.size))


def cosine(x, y) -> float:
    x, y = _f32(x), _f32(y)
    return float(load().orc_cosine_f32(_ptr(x), _ptr(y), x.size))


def l2_subvec(x, y) -> float:
    x, y = _f32(x), _f32(y)
    return float(load().orc_l2_subvec(_ptr(x), _ptr(y), x.size))


def normalize(x) -> np.ndarray:
    x = _f32(x)
    out = np.empty_like(x)
    load().orc_normalize_f32(_ptr(x), x.size, _ptr(out))
    return out


class OracleIndex:
    """Wraps the plain index arrays (lancedb_b200.index.IvfPqIndexData fields)."""

    def __init__(self, *, dim, nlist, m, metric, centroids, codebook, part_offsets, codes_t,
                 row_ids, vectors=None):
        self.centroids = _f32(centroids)
        self.codebook = _f32(codebook)
        self.part_offsets = np.ascontiguousarray(part_offsets, dtype=np.uint64)
        self.codes_t = np.ascontiguousarray(codes_t, dtype=np.uint8)
        self.row_ids = np.ascontiguousarray(row_ids, dtype=np.uint64)
        self.vectors = _f32(vectors) if vectors is not None else None
        self.metric = metric if isinstance(metric, int) else METRICS[metric]
        self.dim, self.nlist, self.m = int(dim), int(nlist), int(m)
        self.c = _Index(self.dim, self.nlist, self.m, self.metric, int(self.row_ids.size),
                        _ptr(self.centroids), _ptr(self.codebook), _ptr(self.part_offsets),
                        _ptr(self.codes_t), _ptr(self.row_ids), _ptr(self.vectors))

    @classmethod
    def from_data(cls, d):
        return cls(dim=d.dim, nlist=d.nlist, m=d.m, metric=d.metric, centroids=d.centroids,
                   codebook=d.codebook, part_offsets=d.part_offsets, codes_t=d.codes_t,
                   row_ids=d.row_ids, vectors=d.vectors)

    def find_partitions(self, q, nprobes):
        q = _f32(q)
        nprobes = min(nprobes, self.nlist)
        parts = np.empty(nprobes, np.uint32)
        dists = np.empty(nprobes, np.float32)
        alld = np.empty(self.nlist, np.float32)
        load().orc_find_partitions(C.byref(self.c), _ptr(q), nprobes, _ptr(parts), _ptr(dists), _ptr(alld))
        return parts, dists, alld

    def build_lut(self, resid_or_query):
        r = _f32(resid_or_query)
        lut = np.empty((self.m, 256), np.float32)
        load().orc_build_lut(C.byref(self.c), _ptr(r), _ptr(lut))
        return lut

    def partition_distances(self, q, part):
        q = _f32(q)
        n = int(self.part_offsets[part + 1] - self.part_offsets[part])
        out = np.empty(n, np.float32)
        load().orc_partition_distances(C.byref(self.c), _ptr(q), int(part), _ptr(out))
        return out

    def ivf_assign(self, vectors):
        v = _f32(vectors).reshape(-1, self.dim)
        out = np.empty(v.shape[0], np.uint32)
        load().orc_ivf_assign(C.byref(self.c), _ptr(v), v.shape[0], _ptr(out))
        return out

    def pq_encode(self, vectors, parts):
        v = _f32(vectors).reshape(-1, self.dim)
        p = np.ascontiguousarray(parts, np.uint32)
        out = np.empty((v.shape[0], self.m), np.uint8)
        load().orc_pq_encode(C.byref(self.c), _ptr(v), _ptr(p), v.shape[0], _ptr(out))
        return out

    def search(self, queries, k=10, nprobes=20, refine_factor=0, lower=None, upper=None, nthreads=1,
               allow=None, allow_bits=0, max_nprobes=0):
        q = _f32(queries).reshape(-1, self.dim)
        B = q.shape[0]
        allow = None if allow is None else np.ascontiguousarray(allow, np.uint32)
        p = _params(k, nprobes, refine_factor, lower, upper, allow, allow_bits, max_nprobes)
        ids = np.empty((B, k), np.uint64)
        dist = np.empty((B, k), np.float32)
        cnt = np.empty(B, np.uint32)
        rc = load().orc_ivfpq_search(C.byref(self.c), _ptr(q), B, C.byref(p), _ptr(ids), _ptr(dist),
                                     _ptr(cnt), nthreads)
        if rc:
            raise RuntimeError(f"orc_ivfpq_search rc={rc}")
        return ids, dist, cnt


def flat_search(vectors, queries, k=10, metric="l2", row_ids=None, lower=None, upper=None, nthreads=1,
                allow=None, allow_bits=0):
    v = _f32(vectors)
    n, dim = v.shape
    q = _f32(queries).reshape(-1, dim)
    B = q.shape[0]
    rid = np.ascontiguousarray(row_ids, dtype=np.uint64) if row_ids is not None else None
    allow = None if allow is None else np.ascontiguousarray(allow, np.uint32)
    p = _params(k, 0, 0, lower, upper, allow, allow_bits)
    ids = np.empty((B, k), np.uint64)
    dist = np.empty((B, k), np.float32)
    cnt = np.empty(B, np.uint32)
    rc = load().orc_flat_search(_ptr(v), n, dim, _ptr(rid), METRICS[metric], _ptr(q), B, C.byref(p),
                                _ptr(ids), _ptr(dist), _ptr(cnt), nthreads)
    if rc:
        raise RuntimeError(f"orc_flat_search rc={rc}")
    return ids, dist, cnt


def pq_scan(lut, codes_t, n, m):
    lut = _f32(lut)
    codes_t = np.ascontiguousarray(codes_t, np.uint8)
    out = np.empty(n, np.float32)
    load().orc_pq_scan(_ptr(lut), _ptr(codes_t), n, m, _ptr(out))
    return out
