"""ctypes bindings used by the tests: the oracle restatement (oracle/librsi_oracle.so, prefix ocl_) and,
when it has been built in this container, the unmodified reference behind oracle/_ref/libref_harness.so
(prefix ref_).  Both expose the same function set so a test can run either through `Lib`."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "librsi_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "rsicnv")
REF_BAMTOOL = os.path.join(ROOT, "oracle", "_ref", "bamtool")

TRANS = {"NBN": 0, "MED": 1, "ALL": 2}


class Cnv(C.Structure):
    """flat mirror of cnv_st (rsi.h:8-51); identical layout in ref_cnv / ocl_cnv / rsigpu_cnv"""
    _fields_ = [(n, C.c_int) for n in ("tid", "type", "geno", "status", "start", "end", "length", "sc1", "sc2", "pair")] + \
               [(n, C.c_double) for n in ("score", "p1", "p2", "cnvmed", "cnvsd", "cnviqr", "refmed", "refsd", "refiqr", "q0")] + \
               [("rp", C.c_int), ("pad_", C.c_int)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "pad_"}

    def key(self):
        return (self.type, self.geno, self.status, self.start, self.end)


def new_cnv(start=0, end=0, type=2, status=0, geno=0, **kw):
    c = Cnv()
    c.tid = -1; c.type = type; c.geno = geno; c.status = status; c.start = start; c.end = end
    c.p1 = 1.0; c.p2 = 1.0; c.q0 = -1.0; c.rp = -1
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Lib:
    def __init__(self, kind: str):
        self.kind = kind
        self.pre = "ocl_" if kind == "oracle" else "ref_"
        self.lib = C.CDLL(ORACLE_SO if kind == "oracle" else REF_SO)
        if kind == "ref":
            self.lib.ref_quiet(1)
        for name in ("median_i32", "median_f32", "median_f64", "iqr_i32", "iqr_f32", "pnorm", "variance_i32",
                     "variance_f32", "apply_cap"):
            self.fn(name).restype = C.c_double
        self.fn("true_median_i32" if kind == "oracle" else "alglib_median_i32").restype = C.c_double

    def fn(self, name):
        return getattr(self.lib, self.pre + name)

    # ---- state
    def set_params(self, m=101, minq=0, min_baseQ=13, cap=4.0, gcadjust=1, trans="NBN", merge=1, threshold=-1.0, epsilon=1.5,
                   chklen=2.5, maxchkbp=100000):
        t = C.c_int(TRANS[trans]) if self.kind == "oracle" else C.c_char_p(trans.encode())
        if self.kind == "ref" and m % 2 != 1:
            m += 1
        self.fn("set_params")(C.c_int(m), C.c_int(minq), C.c_int(min_baseQ), C.c_double(cap), C.c_int(gcadjust), t,
                              C.c_int(merge), C.c_double(threshold), C.c_double(epsilon))
        self.fn("set_knobs")(C.c_double(chklen), C.c_int(maxchkbp))   # -reflen / -maxchkbp (rsi.cpp:2024-2026)

    def set_state(self, RDmedian, RDsd, start, end, Lmax=-1, factor=6.6):
        self.fn("set_state")(C.c_double(RDmedian), C.c_double(RDsd), C.c_int(start), C.c_int(end), C.c_int(Lmax), C.c_double(factor))

    def get_state(self):
        o = np.zeros(10)
        self.fn("get_state")(_p(o, C.c_double))
        return dict(zip(("RDmedian", "RDsd", "start", "end", "Lmax", "factor", "nbnmedian", "nbnlamda", "medmedian", "medlamda"), o.tolist()))

    def set_noncode(self, beg, end):
        b, e = i32(beg), i32(end)
        self.fn("set_noncode")(_p(b, C.c_int), _p(e, C.c_int), C.c_int(len(b)))

    def get_noncode(self):
        b = np.zeros(65536, np.int32); e = np.zeros(65536, np.int32)
        n = self.fn("get_noncode")(_p(b, C.c_int), _p(e, C.c_int), C.c_int(65536))
        return b[:n].copy(), e[:n].copy()

    # ---- L0
    def median(self, x):
        x = np.ascontiguousarray(x)
        if x.dtype == np.int32:
            return self.fn("median_i32")(_p(x, C.c_int), C.c_long(len(x)))
        if x.dtype == np.float32:
            return self.fn("median_f32")(_p(x, C.c_float), C.c_long(len(x)))
        x = x.astype(np.float64)
        return self.fn("median_f64")(_p(x, C.c_double), C.c_long(len(x)))

    def iqr(self, x):
        x = np.ascontiguousarray(x)
        if x.dtype == np.int32:
            return self.fn("iqr_i32")(_p(x, C.c_int), C.c_long(len(x)))
        x = f32(x)
        return self.fn("iqr_f32")(_p(x, C.c_float), C.c_long(len(x)))

    def true_median(self, x):
        x = i32(x)
        return self.fn("true_median_i32" if self.kind == "oracle" else "alglib_median_i32")(_p(x, C.c_int), C.c_long(len(x)))

    def pnorm(self, v):
        return self.fn("pnorm")(C.c_double(v))

    def variance(self, x):
        x = np.ascontiguousarray(x)
        if x.dtype == np.int32:
            return self.fn("variance_i32")(_p(x, C.c_int), C.c_int(len(x)))
        x = f32(x)
        return self.fn("variance_f32")(_p(x, C.c_float), C.c_int(len(x)))

    def runmean(self, y, band):
        y = f32(y); s = np.zeros_like(y)
        self.fn("runmean_f32")(_p(y, C.c_float), _p(s, C.c_float), C.c_int(len(y)), C.c_int(band))
        return s

    # ---- loaders' post-processing
    def noseq_regions(self, fasta):
        fasta = np.ascontiguousarray(fasta, dtype=np.uint8)
        b = np.zeros(65536, np.int32); e = np.zeros(65536, np.int32)
        n = self.fn("noseq_regions")(fasta.ctypes.data_as(C.c_char_p), C.c_int(len(fasta)), _p(b, C.c_int), _p(e, C.c_int), C.c_int(65536))
        return b[:n].copy(), e[:n].copy()

    def checkgccontent(self, rd, gc):
        rd = i32(rd).copy(); gc = np.ascontiguousarray(gc, dtype=np.uint8)
        self.fn("checkgccontent")(_p(rd, C.c_int), _p(gc, C.c_ubyte), C.c_int(len(rd)))
        return rd

    def apply_cap(self, rd):
        rd = i32(rd).copy()
        med = self.fn("apply_cap")(_p(rd, C.c_int), C.c_int(len(rd)))
        return rd, med

    def concatenate(self, rd):
        rd = i32(rd).copy()
        n = self.fn("concatenate")(_p(rd, C.c_int), C.c_int(len(rd)))
        return rd[:n].copy()

    # ---- transforms
    def median_transfer(self, rd, m):
        rd = i32(rd); out = np.zeros(len(rd) // m + 1, np.float32)
        n = self.fn("median_transfer")(_p(rd, C.c_int), C.c_int(len(rd)), C.c_int(m), _p(out, C.c_float))
        return out[:n].copy()

    def nb_transfer(self, rd, m):
        rd = i32(rd); out = np.zeros(len(rd) // m + 1, np.float32)
        n = self.fn("nb_transfer")(_p(rd, C.c_int), C.c_int(len(rd)), C.c_int(m), _p(out, C.c_float))
        return out[:n].copy()

    # ---- RSI scan
    def rsistatus(self, t, medint, tmedian, tlamda, Lmax):
        t = f32(t); medint = i32(medint); st = np.zeros(len(t), np.int32)
        self.fn("rsistatus")(_p(t, C.c_float), _p(medint, C.c_int), C.c_int(len(t)), C.c_double(tmedian), C.c_double(tlamda),
                             C.c_int(Lmax), _p(st, C.c_int))
        return st

    def filterstatus(self, t, dev, status):
        t = f32(t); st = i32(status).copy()
        self.fn("filterstatus")(_p(t, C.c_float), C.c_int(len(t)), C.c_double(dev), _p(st, C.c_int))
        return st

    def _list(self, cap=4096):
        return (Cnv * cap)()

    def continuous_segments(self, status, d=1):
        st = i32(status); out = self._list()
        n = self.fn("continuous_segments")(_p(st, C.c_int), C.c_int(len(st)), C.c_int(d), out, C.c_int(len(out)))
        return [(out[i].start, out[i].end) for i in range(n)]

    def get_rsi_segments(self, t, status, tmedian):
        t = f32(t); st = i32(status); out = self._list()
        n = self.fn("get_rsi_segments")(_p(t, C.c_float), _p(st, C.c_int), C.c_int(len(t)), C.c_double(tmedian), out, C.c_int(len(out)))
        return list(out[:n])

    def rsicnv(self, which, t, medint):
        t = f32(t); medint = i32(medint); st = np.zeros(len(t), np.int32); out = self._list()
        n = self.fn("rsicnv")(C.c_int(which), _p(t, C.c_float), _p(medint, C.c_int), C.c_int(len(t)), _p(st, C.c_int), out, C.c_int(len(out)))
        return st, list(out[:n])

    # ---- candidates
    @staticmethod
    def _arr(lst):
        a = (Cnv * max(len(lst), 1))()
        for i, c in enumerate(lst):
            C.memmove(C.byref(a[i]), C.byref(c), C.sizeof(Cnv))
        return a

    @staticmethod
    def _copy(a, n):
        out = []
        for i in range(n):
            c = Cnv(); C.memmove(C.byref(c), C.byref(a[i]), C.sizeof(Cnv)); out.append(c)
        return out

    def isitcnvwrap(self, rd, lst, idx):
        rd = i32(rd); a = self._arr(lst)
        self.fn("isitcnvwrap")(_p(rd, C.c_int), C.c_int(len(rd)), a, C.c_int(len(lst)), C.c_int(idx))
        return self._copy(a, len(lst))

    def areblockscnv(self, medint, status, lst):
        medint = i32(medint); st = i32(status); a = self._arr(lst)
        n = self.fn("areblockscnv")(_p(medint, C.c_int), _p(st, C.c_int), C.c_int(len(st)), a, C.c_int(len(lst)))
        return self._copy(a, n)

    def sort(self, lst):
        a = self._arr(lst)
        self.fn("sort")(a, C.c_int(len(lst)))
        return self._copy(a, len(lst))

    def optimize(self, rd, lst):
        rd = i32(rd); a = self._arr(lst)
        self.fn("optimize")(_p(rd, C.c_int), C.c_int(len(rd)), a, C.c_int(len(lst)))
        return self._copy(a, len(lst))

    def mergesegments(self, rd, lst):
        rd = i32(rd); a = self._arr(lst)
        n = self.fn("mergesegments")(_p(rd, C.c_int), C.c_int(len(rd)), a, C.c_int(len(lst)))
        return self._copy(a, n)

    def sd_filters(self, lst):
        a = self._arr(lst)
        n = self.fn("sd_filters")(a, C.c_int(len(lst)))
        return self._copy(a, n)

    def expand_coordinate(self, p):
        return self.fn("expand_coordinate")(C.c_int(p))

    def detectcnv(self, rd):
        rd = i32(rd); out = self._list()
        n = self.fn("detectcnv")(_p(rd, C.c_int), C.c_int(len(rd)), out, C.c_int(len(out)))
        return self._copy(out, n)

    def last_list(self, which):
        """intermediate call lists of the last detectcnv (oracle only): 0 segments, 1 blocks, 2 premerge, 3 merged"""
        out = self._list(65536)
        n = self.fn("last_list")(C.c_int(which), out, C.c_int(len(out)))
        return self._copy(out, n)

    def depth_path(self, depth, fasta, stage=3, want_bins=False):
        """returns dict(depth=compacted array, stats=(RDmedian, RDsd), calls=[Cnv], bins=... (oracle only))"""
        depth = i32(depth).copy(); fasta = np.ascontiguousarray(fasta, dtype=np.uint8)
        nc = C.c_int(0); stats = np.zeros(2); out = self._list()
        res = {}
        if self.kind == "oracle":
            nbmax = len(depth) // 3 + 8      # bins of at least 3 bases (m is odd, >= 3)
            bm = np.zeros(nbmax, np.float32) if want_bins else None
            bn = np.zeros(nbmax, np.float32) if want_bins else None
            bi = np.zeros(nbmax, np.int32) if want_bins else None
            bs = np.zeros(nbmax, np.int32) if want_bins else None
            n = self.fn("depth_path")(_p(depth, C.c_int), _p(fasta, C.c_ubyte), C.c_int(len(depth)), C.c_int(stage), C.byref(nc),
                                      _p(stats, C.c_double), out, C.c_int(len(out)),
                                      _p(bm, C.c_float) if want_bins else None, _p(bn, C.c_float) if want_bins else None,
                                      _p(bi, C.c_int) if want_bins else None, _p(bs, C.c_int) if want_bins else None)
            if want_bins:
                res["bins"] = (bm, bn, bi, bs)
        else:
            n = self.fn("depth_path")(_p(depth, C.c_int), fasta.ctypes.data_as(C.c_char_p), C.c_int(len(depth)), C.c_int(stage),
                                      C.byref(nc), _p(stats, C.c_double), out, C.c_int(len(out)))
        res.update(depth=depth[:nc.value].copy(), stats=(stats[0], stats[1]), calls=self._copy(out, n))
        return res

    def format_row(self, c, chrom, rdmedian, rdsd):
        buf = C.create_string_buffer(1024)
        self.fn("format_row")(C.byref(c), C.c_char_p(chrom.encode()), C.c_double(rdmedian), C.c_double(rdsd), buf, C.c_int(1024))
        return buf.value.decode()


def _rp(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def oracle_pileup(o: "Lib", reads: dict, L: int) -> np.ndarray:
    """load_data_from_bam's hot loop on the read SoA (oracle only)"""
    rd = np.zeros(L, np.int32)
    o.lib.ocl_pileup(C.c_int64(len(reads["pos"])), C.c_int(0), _rp(reads["pos"], C.c_int32), _rp(reads["flag"], C.c_uint16),
                     _rp(reads["mapq"], C.c_uint8), _rp(reads["cigar_off"], C.c_uint32), _rp(reads["cigar"], C.c_uint32),
                     _rp(reads["qual_off"], C.c_uint64), _rp(reads["qual"], C.c_uint8), _rp(rd, C.c_int32), C.c_int(L))
    return rd


def oracle_cnv_stat(o: "Lib", reads: dict, L: int, calls: list) -> list:
    """cnv_stat + bam_rd_pr_stats on the read SoA (oracle only); returns the annotated calls"""
    a = Lib._arr(calls)
    o.lib.ocl_cnv_stat(C.c_int64(len(reads["pos"])), C.c_int(0), _rp(reads["pos"], C.c_int32), _rp(reads["mpos"], C.c_int32),
                       _rp(reads["isize"], C.c_int32), _rp(reads["mtid"], C.c_int32), _rp(reads["flag"], C.c_uint16), _rp(reads["mapq"], C.c_uint8),
                       _rp(reads["cigar_off"], C.c_uint32), _rp(reads["cigar"], C.c_uint32), C.c_int(L), a, C.c_int(len(calls)))
    return Lib._copy(a, len(calls))


def oracle_isize(o: "Lib", reads: dict, L: int):
    out = np.zeros(2, np.int32)
    o.lib.ocl_isize_stats(C.c_int64(len(reads["pos"])), C.c_int(0), _rp(reads["pos"], C.c_int32), _rp(reads["mpos"], C.c_int32),
                          _rp(reads["isize"], C.c_int32), _rp(reads["mtid"], C.c_int32), _rp(reads["flag"], C.c_uint16),
                          _rp(reads["cigar_off"], C.c_uint32), _rp(reads["cigar"], C.c_uint32), C.c_int(L), _rp(out, C.c_int32))
    return int(out[0]), int(out[1])


def oracle_bam_path(o: "Lib", reads: dict, fasta, **params):
    """the whole BAM path on the oracle: pileup -> depth path -> sd_filters -> cnv_stat; returns dict(raw, calls, stats)"""
    o.set_params(**params)
    L = len(fasta)
    raw = oracle_pileup(o, reads, L)
    o.set_params(**params)
    res = o.depth_path(raw, fasta, 3)
    calls = oracle_cnv_stat(o, reads, L, res["calls"]) if res["calls"] else []
    return dict(raw=raw, calls=calls, stats=res["stats"])


def have_ref() -> bool:
    return os.path.exists(REF_SO)


# ---- BAM records as the reference's samtools returns them (ref_harness.cpp::ref_bam_records) / as oracle/bam_decode.py restates them
BAM_FIELDS = (("pos", np.int32), ("mpos", np.int32), ("isize", np.int32), ("mtid", np.int32), ("flag", np.uint16), ("mapq", np.uint8),
              ("cigar_off", np.uint32), ("cigar", np.uint32), ("qual_off", np.uint64), ("qual", np.uint8))


def ref_bam_records(path: str, tid: int) -> dict:
    lib = C.CDLL(REF_SO)
    lib.ref_bam_records.restype = C.c_long
    cap_r, cap_c, cap_q = 1 << 16, 1 << 18, 1 << 22
    while True:
        a = {"pos": np.zeros(cap_r, np.int32), "mpos": np.zeros(cap_r, np.int32), "isize": np.zeros(cap_r, np.int32), "mtid": np.zeros(cap_r, np.int32),
             "flag": np.zeros(cap_r, np.uint16), "mapq": np.zeros(cap_r, np.uint8), "cigar_off": np.zeros(cap_r + 1, np.uint32), "cigar": np.zeros(cap_c, np.uint32),
             "qual_off": np.zeros(cap_r + 1, np.uint64), "qual": np.zeros(cap_q, np.uint8)}
        nc = C.c_long(0); nq = C.c_long(0)
        n = lib.ref_bam_records(C.c_char_p(path.encode()), C.c_int(tid), C.c_long(cap_r), C.c_long(cap_c), C.c_long(cap_q),
                                *[C.c_void_p(a[k].ctypes.data) for k, _ in BAM_FIELDS], C.byref(nc), C.byref(nq))
        if n >= 0:
            break
        cap_r = max(cap_r, -1 - n + 1); cap_c = max(cap_c, nc.value + 1); cap_q = max(cap_q, nq.value + 1)
    return {"pos": a["pos"][:n], "mpos": a["mpos"][:n], "isize": a["isize"][:n], "mtid": a["mtid"][:n], "flag": a["flag"][:n], "mapq": a["mapq"][:n],
            "cigar_off": a["cigar_off"][:n + 1], "cigar": a["cigar"][:nc.value], "qual_off": a["qual_off"][:n + 1], "qual": a["qual"][:nq.value]}


def oracle_bam_decode(data) -> dict:
    import importlib.util
    spec = importlib.util.spec_from_file_location("rsi_oracle_bam_decode", os.path.join(ROOT, "oracle", "bam_decode.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.decode(data)
