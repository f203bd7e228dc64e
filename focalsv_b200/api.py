"""ctypes binding of libfocalsv_cuda.so (include/focalsv_cuda.h).

This is the host-side mirror a FocalSV maintainer would call at the sites that
shell out to minimap2 today (DipPAV_variant_call.py:103-112,
call_DUP_from_contigs.py:114-126, align_ins2ref.py:64-71).  It fails loudly:
a missing shared object or a missing CUDA device raises; there is no CPU path.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import RESULT_DTYPE, TASK_DTYPE, Scoring, Stats

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FSV_LIB_PATH") or os.path.join(_HERE, "libfocalsv_cuda.so")   # FSV_LIB_PATH: kernel experiments

# every symbol include/focalsv_cuda.h declares
EXPORTS = ("fsv_init", "fsv_destroy", "fsv_strerror", "fsv_last_error", "fsv_abi_version", "fsv_device_count",
           "fsv_get_stats", "fsv_set_option", "fsv_align_batch", "fsv_batch_create", "fsv_batch_run",
           "fsv_batch_fetch", "fsv_batch_destroy", "fsv_ksw_extz2", "fsv_ksw_extd2", "fsv_task_cells",
           "fsv_lpt_bins", "fsv_measure_int_peak", "fsv_batch_timeline", "fsv_batch_signatures", "fsv_edit_distance_batch",
           "fsv_preset_lookup", "fsv_realign_regions", "fsv_chain_pieces", "fsv_stitch_cigars", "fsv_batch_plan", "fsv_chain_pair")

_lib = None


class FsvError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        RuntimeError.__init__(self, "libfocalsv_cuda error %d: %s" % (code, msg))


def load_library(path=None):
    """Load the CUDA library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FsvError(_abi.ERR_NO_DEVICE, "%s not built: run `python -m focalsv_b200.build` "
                       "(or __graft_entry__.build()); there is no CPU fallback" % p)
    lib = C.CDLL(p)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    lib.fsv_init.argtypes = [C.c_int, C.POINTER(vp)]
    lib.fsv_destroy.argtypes = [vp]
    lib.fsv_destroy.restype = None
    lib.fsv_strerror.argtypes = [C.c_int]
    lib.fsv_strerror.restype = C.c_char_p
    lib.fsv_last_error.argtypes = [vp]
    lib.fsv_last_error.restype = C.c_char_p
    lib.fsv_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.fsv_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.fsv_align_batch.argtypes = [vp, C.POINTER(Scoring), vp, sz, vp, sz, vp, sz, vp, vp, sz, C.POINTER(sz)]
    lib.fsv_batch_create.argtypes = [vp, C.POINTER(Scoring), vp, sz, vp, sz, vp, sz, C.POINTER(vp)]
    lib.fsv_batch_run.argtypes = [vp]
    lib.fsv_batch_fetch.argtypes = [vp, vp, vp, sz, C.POINTER(sz)]
    lib.fsv_batch_destroy.argtypes = [vp]
    lib.fsv_batch_timeline.argtypes = [vp, vp]
    lib.fsv_batch_plan.argtypes = [vp, vp]
    i8p = C.POINTER(C.c_int8)
    lib.fsv_ksw_extz2.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int8, i8p, C.c_int8, C.c_int8, C.c_int, C.c_int, C.c_int, C.c_int,
                                  vp, vp, C.c_int]
    lib.fsv_ksw_extd2.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int8, i8p, C.c_int8, C.c_int8, C.c_int8, C.c_int8, C.c_int, C.c_int,
                                  C.c_int, C.c_int, vp, vp, C.c_int]
    lib.fsv_batch_signatures.argtypes = [vp, vp, C.c_int, vp, sz, C.POINTER(sz)]
    lib.fsv_edit_distance_batch.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp]
    lib.fsv_batch_destroy.restype = None
    lib.fsv_chain_pieces.argtypes = [vp, i32, vp, i32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, sz, C.POINTER(sz),
                                     C.POINTER(i32), C.POINTER(i32)]
    lib.fsv_stitch_cigars.argtypes = [vp, vp, sz, vp, vp, vp, sz, C.POINTER(sz)]
    lib.fsv_chain_pair.argtypes = [vp, i32, vp, i32, C.POINTER(_abi.ChainOpts), vp, sz, C.POINTER(sz), vp, sz, C.POINTER(sz)]
    lib.fsv_preset_lookup.argtypes = [C.c_char_p, C.POINTER(_abi.PresetC), C.POINTER(Scoring)]
    lib.fsv_realign_regions.argtypes = [vp, vp, sz, vp, vp, vp, sz, vp, vp, sz, C.c_char_p, C.c_int, C.c_int, vp, vp, sz, C.POINTER(sz)]
    lib.fsv_task_cells.argtypes = [i32, i32, i32]
    lib.fsv_task_cells.restype = i64
    lib.fsv_lpt_bins.argtypes = [vp, sz, C.c_int, vp]
    lib.fsv_measure_int_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    if lib.fsv_abi_version() != _abi.ABI_VERSION:
        raise FsvError(_abi.ERR_INVALID, "ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def preset_lookup(name):
    """(fields dict, Scoring) of a preset as the library knows it (fsv_preset_lookup; host only)."""
    lib = load_library()
    p = _abi.PresetC(); sc = Scoring()
    rc = lib.fsv_preset_lookup(name.encode(), C.byref(p), C.byref(sc))
    if rc != 0:
        raise FsvError(rc, "unknown preset %r" % name)
    return {k: getattr(p, k) for k, _ in _abi.PresetC._fields_[1:]}, sc


def chain_pieces(query, target, k=19, w=19, max_occ=50, max_gap=100000, min_fill=200):
    """fsv_chain_pieces (host only): minimizer seeding + chaining + decomposition of one (query, target) pair of code
    arrays into the DP pieces minimap2's mm_align1 would fill.  Returns (pieces[PIECE_DTYPE], chain_score, n_anchors)."""
    lib = load_library()
    q = np.ascontiguousarray(query, dtype=np.uint8); t = np.ascontiguousarray(target, dtype=np.uint8)
    cap = 1024
    while True:
        out = np.zeros(cap, dtype=_abi.PIECE_DTYPE)
        n = C.c_size_t(0); sc = C.c_int32(0); na = C.c_int32(0)
        rc = lib.fsv_chain_pieces(q.ctypes.data, q.size, t.ctypes.data, t.size, k, w, max_occ, max_gap, min_fill,
                                  out.ctypes.data, cap, C.byref(n), C.byref(sc), C.byref(na))
        if rc == _abi.ERR_CIGAR_CAP:
            cap = int(n.value) + 16
            continue
        if rc != 0:
            raise FsvError(rc, lib.fsv_strerror(rc).decode())
        return out[:n.value], sc.value, na.value


def chain_pair(query, target, opts):
    """fsv_chain_pair (host only): both strands, primary + supplementary chains, extension offers.
    Returns (chains[CHAIN_DTYPE], pieces[PIECE_DTYPE])."""
    lib = load_library()
    q = np.ascontiguousarray(query, dtype=np.uint8); t = np.ascontiguousarray(target, dtype=np.uint8)
    ccap, pcap = 8, 1024
    while True:
        ch = np.zeros(ccap, dtype=_abi.CHAIN_DTYPE); pc = np.zeros(pcap, dtype=_abi.PIECE_DTYPE)
        nc = C.c_size_t(0); npc = C.c_size_t(0)
        rc = lib.fsv_chain_pair(q.ctypes.data, q.size, t.ctypes.data, t.size, C.byref(opts), ch.ctypes.data, ccap, C.byref(nc),
                                pc.ctypes.data, pcap, C.byref(npc))
        if rc == _abi.ERR_CIGAR_CAP:
            ccap, pcap = int(nc.value) + 4, int(npc.value) + 16
            continue
        if rc != 0:
            raise FsvError(rc, lib.fsv_strerror(rc).decode())
        return ch[:nc.value], pc[:npc.value]


def stitch_cigars(pieces, task_of, res, cigar_arena):
    """fsv_stitch_cigars (host only): one CIGAR (uint32 BAM words) from the pieces of one pair."""
    lib = load_library()
    pieces = np.ascontiguousarray(pieces, dtype=_abi.PIECE_DTYPE); task_of = np.ascontiguousarray(task_of, dtype=np.int32)
    res = np.ascontiguousarray(res, dtype=RESULT_DTYPE); arena = np.ascontiguousarray(cigar_arena, dtype=np.uint32)
    cap = 256
    while True:
        out = np.zeros(cap, dtype=np.uint32)
        n = C.c_size_t(0)
        rc = lib.fsv_stitch_cigars(pieces.ctypes.data, task_of.ctypes.data, len(pieces), res.ctypes.data, arena.ctypes.data,
                                   out.ctypes.data, cap, C.byref(n))
        if rc == _abi.ERR_CIGAR_CAP:
            cap = int(n.value) + 16
            continue
        if rc != 0:
            raise FsvError(rc, lib.fsv_strerror(rc).decode())
        return out[:n.value]


def task_cells(qlen, tlen, w):
    return int(load_library().fsv_task_cells(qlen, tlen, w))


def lpt_bins(tasks, n_bins):
    """Length-balanced bins of independent tasks: returns bin index per task."""
    tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
    out = np.zeros(len(tasks), dtype=np.int32)
    rc = load_library().fsv_lpt_bins(tasks.ctypes.data, len(tasks), int(n_bins), out.ctypes.data)
    if rc != 0:
        raise FsvError(rc, "fsv_lpt_bins")
    return out


def make_tasks(qlens, tlens, w, zdrop, end_bonus=0, flag=0, q_offs=None, t_offs=None):
    """Task table for sequences stored back to back in two arenas."""
    qlens = np.asarray(qlens, dtype=np.int64)
    tlens = np.asarray(tlens, dtype=np.int64)
    n = len(qlens)
    t = np.zeros(n, dtype=TASK_DTYPE)
    t["qlen"], t["tlen"] = qlens, tlens
    t["q_off"] = np.concatenate([[0], np.cumsum(qlens)[:-1]]) if q_offs is None else q_offs
    t["t_off"] = np.concatenate([[0], np.cumsum(tlens)[:-1]]) if t_offs is None else t_offs
    t["w"], t["zdrop"], t["end_bonus"], t["flag"] = w, zdrop, end_bonus, flag
    return t


class Aligner(object):
    """One context per device per process (the reference forks joblib workers; create it lazily in each)."""

    def __init__(self, device=-1):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.fsv_init(int(device), C.byref(h))
        if rc != 0:
            raise FsvError(rc, self._lib.fsv_strerror(rc).decode())
        self._h = h
        for kv in filter(None, os.environ.get("FSV_SET", "").split(",")):      # experiments: FSV_SET=key=value,key=value
            k, v = kv.split("=")
            self.set_option(k.strip(), int(v))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fsv_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise FsvError(rc, "%s: %s (%s)" % (what, self._lib.fsv_strerror(rc).decode(),
                                                self._lib.fsv_last_error(self._h).decode()))

    def set_option(self, key, value):
        self._check(self._lib.fsv_set_option(self._h, key.encode(), int(value)), "fsv_set_option")

    def stats(self):
        s = Stats()
        self._check(self._lib.fsv_get_stats(self._h, C.byref(s)), "fsv_get_stats")
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def int_peak(self, kind=5):
        """Measured integer issue rate (lane-ops/s) of this device; see fsv_measure_int_peak."""
        v = C.c_double(0)
        self._check(self._lib.fsv_measure_int_peak(self._h, int(kind), C.byref(v)), "fsv_measure_int_peak")
        return v.value

    @staticmethod
    def _arena(a):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        return a

    def align_batch(self, sc, qarena, tarena, tasks, cigar_cap=None, out=None, cig=None):
        """Host buffers in, host buffers out (H2D + kernels + D2H).  Returns (results, cigar_arena).
        `out` / `cig` may be caller-provided (e.g. pinned) buffers, as in the C ABI."""
        qarena, tarena = self._arena(qarena), self._arena(tarena)
        tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
        n = len(tasks)
        if out is None:
            out = np.zeros(n, dtype=RESULT_DTYPE)
        if cig is not None:
            cigar_cap = len(cig)
        if cigar_cap is None:
            with_c = (tasks["flag"] & _abi.EZ_SCORE_ONLY) == 0
            cigar_cap = int((tasks["qlen"].astype(np.int64) + tasks["tlen"] + 2)[with_c].sum()) + 16
        if cig is None:
            cig = np.zeros(max(int(cigar_cap), 1), dtype=np.uint32)
        used = C.c_size_t(0)
        rc = self._lib.fsv_align_batch(self._h, C.byref(sc), qarena.ctypes.data, qarena.size, tarena.ctypes.data,
                                       tarena.size, tasks.ctypes.data, n, out.ctypes.data, cig.ctypes.data,
                                       int(cigar_cap), C.byref(used))
        if rc == _abi.ERR_CIGAR_CAP:
            raise FsvError(rc, "cigar arena too small: need %d words" % used.value)
        self._check(rc, "fsv_align_batch")
        return out, cig[:used.value]

    def realign_regions(self, ref_codes, region_start, region_end, contig_codes, contig_off, contig_len,
                        preset="asm5", bw=2000, flag=0, cigar_cap=None):
        """fsv_realign_regions: contig i against ref_codes[region_start[i]:region_end[i]] with the preset's scoring
        (the step around DipPAV_variant_call.py:103's minimap2 call).  Returns (records, cigar_arena)."""
        ref = self._arena(ref_codes); ctg = self._arena(contig_codes)
        rs = np.ascontiguousarray(region_start, dtype=np.int64); re_ = np.ascontiguousarray(region_end, dtype=np.int64)
        co = np.ascontiguousarray(contig_off, dtype=np.int64); cl = np.ascontiguousarray(contig_len, dtype=np.int32)
        n = len(rs)
        if not (len(re_) == len(co) == len(cl) == n):
            raise ValueError("realign_regions: one region and one contig per pair")
        if cigar_cap is None:
            cigar_cap = int((re_ - rs).sum() + cl.astype(np.int64).sum()) + 2 * n + 16
        rec = np.zeros(max(n, 1), dtype=_abi.RECORD_DTYPE)
        cig = np.zeros(max(int(cigar_cap), 1), dtype=np.uint32)
        used = C.c_size_t(0)
        rc = self._lib.fsv_realign_regions(self._h, ref.ctypes.data, ref.size, rs.ctypes.data, re_.ctypes.data, ctg.ctypes.data,
                                           ctg.size, co.ctypes.data, cl.ctypes.data, n, preset.encode(), int(bw), int(flag),
                                           rec.ctypes.data, cig.ctypes.data, int(cigar_cap), C.byref(used))
        if rc == _abi.ERR_CIGAR_CAP:
            raise FsvError(rc, "cigar arena too small: need %d words" % used.value)
        self._check(rc, "fsv_realign_regions")
        return rec[:n], cig[:used.value]

    def edit_distances(self, seqs_a, seqs_b):
        """Global unit-cost edit distance of seqs_a[i] vs seqs_b[i] (what the reference asks edlib for,
        remove_redundancy.py:57-63).  Sequences: str, bytes or uint8 arrays.  Returns an int32 array."""
        def arena(seqs):
            bufs = [np.frombuffer(s.encode() if isinstance(s, str) else bytes(s), dtype=np.uint8) if not isinstance(s, np.ndarray)
                    else np.ascontiguousarray(s, dtype=np.uint8) for s in seqs]
            lens = np.array([len(x) for x in bufs], dtype=np.int64)
            offs = np.concatenate([[0], np.cumsum(lens)[:-1]]) if len(bufs) else np.zeros(0, np.int64)
            return (np.concatenate(bufs) if len(bufs) and lens.sum() else np.zeros(1, np.uint8)), offs, lens
        if len(seqs_a) != len(seqs_b):
            raise ValueError("edit_distances needs as many a as b sequences")
        a, ao, al_ = arena(seqs_a)
        b, bo, bl = arena(seqs_b)
        n = len(al_)
        pairs = np.zeros(n, dtype=_abi.PAIR_DTYPE)
        pairs["a_off"], pairs["b_off"], pairs["a_len"], pairs["b_len"] = ao, bo, al_, bl
        out = np.zeros(max(n, 1), dtype=np.int32)
        rc = self._lib.fsv_edit_distance_batch(self._h, a.ctypes.data, int(al_.sum()), b.ctypes.data, int(bl.sum()),
                                               pairs.ctypes.data, n, out.ctypes.data)
        self._check(rc, "fsv_edit_distance_batch")
        return out[:n]

    def batch(self, sc, qarena, tarena, tasks):
        return Batch(self, sc, qarena, tarena, tasks)

    def _single(self, dual, query, target, sc, w, zdrop, end_bonus, flag):
        query, target = self._arena(query), self._arena(target)
        tasks = make_tasks([len(query)], [len(target)], w, zdrop, end_bonus, flag)
        if dual != (sc.q2 >= 0):
            raise FsvError(_abi.ERR_INVALID, "scoring does not match the requested gap model")
        res, cig = self.align_batch(sc, query, target, tasks)
        return res[0], cig

    def extz2(self, query, target, sc, w=-1, zdrop=-1, end_bonus=0, flag=0):
        """ksw_extz2_sse (ksw2.h:54-55) on the GPU."""
        return self._single(False, query, target, sc, w, zdrop, end_bonus, flag)

    def extd2(self, query, target, sc, w=-1, zdrop=-1, end_bonus=0, flag=0):
        """ksw_extd2_sse (ksw2.h:60-61) on the GPU."""
        return self._single(True, query, target, sc, w, zdrop, end_bonus, flag)


class Batch(object):
    """Staged batch: create = H2D, run = kernels on HBM-resident inputs, fetch = D2H."""

    def __init__(self, aligner, sc, qarena, tarena, tasks):
        self._al = aligner
        self._lib = aligner._lib
        self.qarena, self.tarena = aligner._arena(qarena), aligner._arena(tarena)
        self.tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
        self.sc = sc
        h = C.c_void_p()
        rc = self._lib.fsv_batch_create(aligner._h, C.byref(sc), self.qarena.ctypes.data, self.qarena.size,
                                        self.tarena.ctypes.data, self.tarena.size, self.tasks.ctypes.data,
                                        len(self.tasks), C.byref(h))
        aligner._check(rc, "fsv_batch_create")
        self._h = h

    def run(self):
        self._al._check(self._lib.fsv_batch_run(self._h), "fsv_batch_run")

    def fetch(self, cigar_cap=None):
        n = len(self.tasks)
        out = np.zeros(n, dtype=RESULT_DTYPE)
        if cigar_cap is None:
            cigar_cap = int((self.tasks["qlen"].astype(np.int64) + self.tasks["tlen"] + 2).sum()) + 16
        cig = np.zeros(max(int(cigar_cap), 1), dtype=np.uint32)
        used = C.c_size_t(0)
        rc = self._lib.fsv_batch_fetch(self._h, out.ctypes.data, cig.ctypes.data, int(cigar_cap), C.byref(used))
        self._al._check(rc, "fsv_batch_fetch")
        return out, cig[:used.value]

    def signatures(self, ref_start=None, min_svlen=30):
        """DEL / INS signatures of every task, extracted on the device from the CIGARs of the last run
        (extract_contig_signature_CCS.py:14-127); only these records are copied back.  `ref_start[i]` is the
        reference coordinate of target[0] of task i.  Returns a SIGNATURE_DTYPE array (per task: DELs, then INSs)."""
        n = len(self.tasks)
        rs = None if ref_start is None else np.ascontiguousarray(ref_start, dtype=np.int64)
        if rs is not None and len(rs) != n:
            raise ValueError("ref_start must have one entry per task")
        cap = 1024
        while True:
            out = np.zeros(cap, dtype=_abi.SIGNATURE_DTYPE)
            used = C.c_size_t(0)
            rc = self._lib.fsv_batch_signatures(self._h, None if rs is None else rs.ctypes.data, int(min_svlen),
                                                out.ctypes.data, cap, C.byref(used))
            if rc == _abi.ERR_CIGAR_CAP:
                cap = int(used.value) + 16
                continue
            self._al._check(rc, "fsv_batch_signatures")
            return out[:used.value]

    def timeline(self):
        """(n, 2) int64: device start / end time (ns) of every task in the last run."""
        out = np.zeros((len(self.tasks), 2), dtype=np.int64)
        self._al._check(self._lib.fsv_batch_timeline(self._h, out.ctypes.data), "fsv_batch_timeline")
        return out

    def plan(self):
        """Per task: _abi.PLAN_* bits (kernel family, segmented, exclusive launch), warps in bits 8..11, segments in bits 16.. ."""
        out = np.zeros(max(len(self.tasks), 1), dtype=np.int32)
        self._al._check(self._lib.fsv_batch_plan(self._h, out.ctypes.data), "fsv_batch_plan")
        return out[:len(self.tasks)]

    def close(self):
        # (a batch that outlives its Aligner was detached by fsv_destroy: destroying it only frees the host object)
        if getattr(self, "_h", None):
            self._lib.fsv_batch_destroy(self._h)
            self._h = None

    __del__ = close


def task_cigar(res_row, cigar_arena):
    o, n = int(res_row["cigar_off"]), int(res_row["n_cigar"])
    return cigar_arena[o:o + n]
