"""rabbitvar_b200 — B200-native pileup-and-score path of RabbitVar.

Thin ctypes binding over the C ABI in ``include/rabbitvar_b200.h`` / ``include/rabbitvar_b200_host.h``
(``librvgpu.so``, built in-tree by ``__graft_entry__.build()``).  The names mirror the reference's
per-region stage calls (``CigarParser::process`` -> :meth:`Context.pileup`,
``ToVarsBuilder::process`` -> :meth:`Context.score`, ``one_region_run`` -> :meth:`Context.call_regions`).

There is no CPU fallback: importing works anywhere (symbol checks), every compute call raises
``RabbitVarError`` when the CUDA library or a GPU is missing.
"""
import ctypes as C
import os

__all__ = ["RabbitVarError", "lib", "lib_path", "Params", "Limits", "Read", "Region", "Event", "Variant",
           "PileupStats", "Timing", "Context", "Pipeline", "HostBatch", "default_params", "default_limits", "fetch_ref"]

_HERE = os.path.dirname(os.path.abspath(__file__))


class RabbitVarError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "librvgpu.so")


_lib = None


def lib():
    """Loads librvgpu.so (raises loudly when it has not been built)."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise RabbitVarError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(the CUDA extension is required; there is no CPU path)")
        _lib = C.CDLL(p)
        _declare(_lib)
    return _lib


class Params(C.Structure):
    _fields_ = [("goodq", C.c_double), ("freq", C.c_double), ("lofreq", C.c_double), ("qratio", C.c_double),
                ("mapq", C.c_double), ("bias", C.c_double), ("vext", C.c_int32), ("mismatch", C.c_int32),
                ("minr", C.c_int32), ("min_bias_reads", C.c_int32), ("read_pos_filter", C.c_int32),
                ("minmatch", C.c_int32), ("trim_bases_after", C.c_int32), ("indelsize", C.c_int32),
                ("mapping_quality", C.c_int32), ("samfilter", C.c_int32), ("local_realign", C.c_uint8),
                ("move3", C.c_uint8), ("uniq_u", C.c_uint8), ("uniq_un", C.c_uint8), ("dedup", C.c_uint8),
                ("pileup", C.c_uint8), ("fisher", C.c_uint8), ("has_bam2", C.c_uint8), ("candidates_only", C.c_uint8),
                ("pad_", C.c_uint8 * 7)]


class Limits(C.Structure):
    _fields_ = [("max_reads", C.c_int64), ("max_read_bytes", C.c_int64), ("max_positions", C.c_int64),
                ("max_regions", C.c_int32), ("halo", C.c_int32), ("max_events", C.c_int64),
                ("max_variants", C.c_int64), ("max_patch", C.c_int64), ("max_ref_bases", C.c_int64),
                ("max_sparse_obs", C.c_int64)]


class Read(C.Structure):
    _fields_ = [("pos", C.c_int32), ("mpos", C.c_int32), ("data_off16", C.c_uint32), ("l_seq", C.c_int32),
                ("flag", C.c_uint16), ("n_cigar", C.c_uint16), ("nm", C.c_int16), ("mapq", C.c_uint8),
                ("mate_same_tid", C.c_uint8), ("end_pos", C.c_int32), ("mtid", C.c_int32)]


class ReadBatch(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("reads", C.c_void_p), ("pool", C.c_void_p), ("pool_bytes", C.c_int64)]


class Region(C.Structure):
    _fields_ = [("start", C.c_int32), ("end", C.c_int32), ("ref_lo", C.c_int32), ("ref_hi", C.c_int32),
                ("read_lo", C.c_int64), ("read_hi", C.c_int64), ("chr_len", C.c_int32),
                ("max_read_len_in", C.c_int32)]


class Event(C.Structure):
    _fields_ = [("region", C.c_int32), ("pos", C.c_int32), ("read_idx", C.c_uint32), ("seq_no", C.c_uint16),
                ("kind", C.c_uint8), ("flags", C.c_uint8), ("tp", C.c_int32), ("nm", C.c_int32),
                ("qsum", C.c_int32), ("qcnt", C.c_int32), ("dir", C.c_uint8), ("mapq", C.c_uint8),
                ("keylen", C.c_uint8), ("pad", C.c_uint8), ("aux0", C.c_int32), ("aux1", C.c_int32),
                ("aux2", C.c_int32), ("key", C.c_char * 112)]


class Variant(C.Structure):
    _fields_ = [("region", C.c_int32), ("pos", C.c_int32), ("cnt", C.c_int32), ("fwd", C.c_int32),
                ("rev", C.c_int32), ("tcov", C.c_int32), ("hicnt", C.c_int32), ("hicov", C.c_int32),
                ("ref_fwd", C.c_int32), ("ref_rev", C.c_int32), ("shift3", C.c_int32), ("msint", C.c_int32),
                ("freq", C.c_double), ("pmean", C.c_double), ("qual", C.c_double), ("mapq", C.c_double),
                ("qratio", C.c_double), ("hifreq", C.c_double), ("extrafreq", C.c_double), ("nm", C.c_double),
                ("msi", C.c_double), ("pvalue", C.c_double), ("oddratio", C.c_double), ("bias_ref", C.c_uint8),
                ("bias_var", C.c_uint8), ("pstd", C.c_uint8), ("qstd", C.c_uint8), ("is_ref", C.c_uint8),
                ("key_kind", C.c_uint8), ("rank", C.c_uint8), ("pad", C.c_uint8), ("key_id", C.c_int32)]


class PileupStats(C.Structure):
    _fields_ = [("n_items", C.c_int64), ("n_reads_kept", C.c_int64), ("n_aligned_bases", C.c_int64),
                ("n_events", C.c_int64), ("n_overflow", C.c_int64), ("n_unsupported", C.c_int64),
                ("n_walk_items", C.c_int64), ("n_walk_full", C.c_int64), ("n_clipped", C.c_int64),
                ("n_score_unsupported", C.c_int64), ("n_sparse_obs", C.c_int64), ("n_walk_segments", C.c_int64)]


class Timing(C.Structure):
    _fields_ = [("push_ms", C.c_double), ("pileup_ms", C.c_double), ("fetch_ms", C.c_double), ("host_ms", C.c_double),
                ("patch_ms", C.c_double), ("score_ms", C.c_double), ("assemble_ms", C.c_double),
                ("pileup_kernel_ms", C.c_float), ("score_kernel_ms", C.c_float), ("n_items", C.c_int64),
                ("n_reads_kept", C.c_int64), ("n_aligned_bases", C.c_int64), ("n_events", C.c_int64),
                ("n_unsupported", C.c_int64), ("n_variants", C.c_int64), ("n_lines", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


# every symbol include/*.h declares (checked by tests without a GPU)
ABI_SYMBOLS = [
    "rv_abi_version", "rv_device_count", "rv_warmup", "rv_default_params", "rv_default_limits", "rv_create", "rv_destroy",
    "rv_last_error", "rv_sync", "rv_set_lazy", "rv_ctx_halo", "rv_set_params", "rv_set_reference", "rv_push_reads", "rv_push_reads_range", "rv_push_reads_ranges", "rv_push_reads_device", "rv_set_regions",
    "rv_pileup", "rv_score", "rv_score_positions", "rv_get_pileup_stats", "rv_fetch_max_read_len", "rv_fetch_tables", "rv_fetch_rows",
    "rv_fetch_events",
    "rv_apply_patch", "rv_fetch_variants", "rv_cov_summary", "rv_variant_count", "rv_fisher_exact", "rv_last_kernel_ms", "rv_last_pileup_split_ms", "rv_last_pileup_stage_ms", "rv_timer_start",
    "rv_timer_stop", "rv_launch_count",
    "rvh_load_bam", "rvh_batch_append", "rvh_batch_n_reads", "rvh_batch_reads", "rvh_batch_pool",
    "rvh_batch_pool_bytes", "rvh_batch_max_ref_span", "rvh_batch_pin", "rvh_batch_free", "rvh_make_regions", "rvh_fetch_ref",
    "rvh_call_regions", "rvh_install_patch", "rvh_last_error",
    "rvh_pipeline_create", "rvh_pipeline_destroy", "rvh_pipeline_run", "rvh_pipeline_run_paired", "rvh_pipeline_launch_count",
    "rvh_inflate_block", "rvh_crc32", "rvh_run_files", "rvh_free",
]


def _declare(L):
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.rv_abi_version.restype = C.c_int
    L.rv_device_count.restype = C.c_int
    L.rvh_inflate_block.argtypes = [vp, i64, vp, i64]
    L.rvh_inflate_block.restype = i64
    L.rvh_crc32.argtypes = [vp, i64]
    L.rvh_crc32.restype = C.c_uint32
    L.rvh_run_files.argtypes = [C.POINTER(Params), C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, i32, C.POINTER(C.c_char_p),
                                C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_char_p), i32, i32, i32, C.POINTER(vp),
                                C.POINTER(i64), C.POINTER(C.c_double)]
    L.rvh_free.argtypes = [vp]
    L.rv_default_params.argtypes = [C.POINTER(Params)]
    L.rv_default_limits.argtypes = [C.POINTER(Limits)]
    L.rv_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Params), C.POINTER(Limits)]
    L.rv_destroy.argtypes = [vp]
    L.rv_last_error.argtypes = [vp]
    L.rv_last_error.restype = C.c_char_p
    L.rv_sync.argtypes = [vp]
    L.rv_ctx_halo.argtypes = [vp]
    L.rv_ctx_halo.restype = i32
    L.rv_set_params.argtypes = [vp, C.POINTER(Params)]
    L.rv_set_reference.argtypes = [vp, i32, i64, vp]
    L.rv_push_reads.argtypes = [vp, C.POINTER(ReadBatch)]
    L.rv_push_reads_device.argtypes = [vp, C.POINTER(ReadBatch)]
    L.rv_push_reads_range.argtypes = [vp, C.POINTER(ReadBatch), i64, i64]
    L.rv_push_reads_ranges.argtypes = [vp, C.POINTER(ReadBatch), i32, vp, vp]
    L.rvh_pipeline_create.argtypes = [C.c_int, C.c_int]
    L.rvh_pipeline_create.restype = vp
    L.rvh_pipeline_destroy.argtypes = [vp]
    L.rvh_pipeline_run.argtypes = [vp, C.POINTER(Params), vp, C.POINTER(Region), i32, i32, vp, i32, i64, C.c_char_p,
                                   C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(i64), C.POINTER(Timing)]
    L.rvh_pipeline_run_paired.argtypes = L.rvh_pipeline_run.argtypes
    L.rvh_pipeline_launch_count.argtypes = [vp]
    L.rvh_pipeline_launch_count.restype = i64
    L.rv_set_regions.argtypes = [vp, C.POINTER(Region), i32]
    L.rv_pileup.argtypes = [vp]
    L.rv_score.argtypes = [vp]
    L.rv_score_positions.argtypes = [vp, vp, vp, i64]
    L.rv_get_pileup_stats.argtypes = [vp, C.POINTER(PileupStats)]
    L.rv_fetch_max_read_len.argtypes = [vp, C.POINTER(C.POINTER(i32)), C.POINTER(i32)]
    L.rv_fetch_tables.argtypes = [vp, i32, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_uint32)),
                                  C.POINTER(i32), C.POINTER(i32)]
    L.rv_fetch_rows.argtypes = [vp, vp, vp, i64, C.POINTER(C.POINTER(C.c_uint32))]
    L.rv_fetch_events.argtypes = [vp, C.POINTER(C.POINTER(Event)), C.POINTER(i64)]
    L.rv_apply_patch.argtypes = [vp, vp, i64, vp, vp, vp, i64]
    L.rv_fetch_variants.argtypes = [vp, C.POINTER(C.POINTER(Variant)), C.POINTER(i64)]
    L.rv_fisher_exact.argtypes = [vp, vp, i64, vp]
    L.rv_cov_summary.argtypes = [vp, vp, vp]
    L.rv_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.rv_last_pileup_split_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.rv_last_pileup_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.rv_timer_start.argtypes = [vp]
    L.rv_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.rv_launch_count.argtypes = [vp]
    L.rv_launch_count.restype = i64
    L.rvh_load_bam.argtypes = [C.c_char_p, C.c_char_p, i32, i32, C.POINTER(i32)]
    L.rvh_load_bam.restype = vp
    L.rvh_batch_append.argtypes = [vp, vp]
    L.rvh_batch_append.restype = i64
    L.rvh_batch_n_reads.argtypes = [vp]
    L.rvh_batch_n_reads.restype = i64
    L.rvh_batch_reads.argtypes = [vp]
    L.rvh_batch_reads.restype = vp
    L.rvh_batch_pool.argtypes = [vp]
    L.rvh_batch_pool.restype = vp
    L.rvh_batch_pool_bytes.argtypes = [vp]
    L.rvh_batch_pool_bytes.restype = i64
    L.rvh_batch_max_ref_span.argtypes = [vp]
    L.rvh_batch_max_ref_span.restype = i32
    L.rvh_batch_pin.argtypes = [vp]
    L.rvh_batch_free.argtypes = [vp]
    L.rvh_make_regions.argtypes = [vp, vp, vp, i32, i32, i32, i64, i64, C.POINTER(Region)]
    L.rvh_fetch_ref.argtypes = [C.c_char_p, C.c_char_p, i32, i32, vp]
    L.rvh_fetch_ref.restype = i64
    L.rvh_call_regions.argtypes = [vp, C.POINTER(Params), vp, C.POINTER(Region), i32, vp, i32, i64, C.c_int,
                                   C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(i64), C.POINTER(Timing)]
    L.rvh_install_patch.argtypes = [vp, C.POINTER(Params), vp, C.POINTER(Region), i32, vp, i32, i64]
    L.rv_variant_count.argtypes = [vp]
    L.rv_variant_count.restype = i64
    L.rvh_last_error.restype = C.c_char_p


class Pipeline:
    """Chunked, multi-worker form of :meth:`Context.call_regions` (``rvh_pipeline_*``): several contexts on one
    GPU work through the region list chunk by chunk so copies, kernels and the host stages overlap."""

    def __init__(self, device=0, n_workers=4):
        if lib().rv_device_count() <= 0:
            raise RabbitVarError("no CUDA device: the pipeline has no CPU path")
        self._h = lib().rvh_pipeline_create(device, n_workers)
        if not self._h:
            raise RabbitVarError("rvh_pipeline_create failed")

    def run(self, params, batch, regions, chunk_regions, ref_bases, ref_lo, sample, chrom, paired=False, raw=False):
        """paired=True: regions = n tumor tiles followed by the same n normal tiles; returns the somatic-mode TSV.
        raw=True returns a memoryview of the library-owned text (valid until the next run) instead of a str copy."""
        out = C.c_char_p()
        n = C.c_int64()
        tm = Timing()
        fn = lib().rvh_pipeline_run_paired if paired else lib().rvh_pipeline_run
        rc = fn(self._h, C.byref(params), batch._h, regions, len(regions), chunk_regions,
                                    C.cast(C.c_char_p(ref_bases), C.c_void_p), ref_lo, len(ref_bases), sample.encode(),
                                    chrom.encode(), C.byref(out), C.byref(n), C.byref(tm))
        if rc != 0:
            raise RabbitVarError(f"rvh_pipeline_run failed ({rc}): {lib().rvh_last_error().decode()}")
        if raw:
            addr = C.cast(out, C.c_void_p).value
            return memoryview((C.c_char * n.value).from_address(addr) if n.value else b""), tm
        return C.string_at(out, n.value).decode(), tm

    def launch_count(self):
        return lib().rvh_pipeline_launch_count(self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib().rvh_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_params(**kw):
    p = Params()
    lib().rv_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_limits(**kw):
    l = Limits()
    lib().rv_default_limits(C.byref(l))
    for k, v in kw.items():
        setattr(l, k, v)
    return l


class HostBatch:
    """Decoded reads of one BAM span in the staging layout (host memory)."""

    def __init__(self, bam_path, chrom, start, end):
        clen = C.c_int32(0)
        self._h = lib().rvh_load_bam(os.fsencode(bam_path), chrom.encode(), int(start), int(end), C.byref(clen))
        if not self._h:
            raise RabbitVarError("rvh_load_bam: " + lib().rvh_last_error().decode())
        self.chr_len = clen.value
        self.chrom = chrom

    def append(self, other):
        return lib().rvh_batch_append(self._h, other._h)

    @property
    def n_reads(self):
        return lib().rvh_batch_n_reads(self._h)

    @property
    def pool_bytes(self):
        return lib().rvh_batch_pool_bytes(self._h)

    @property
    def reads_ptr(self):
        return lib().rvh_batch_reads(self._h)

    @property
    def pool_ptr(self):
        return lib().rvh_batch_pool(self._h)

    def reads_numpy(self):
        import numpy as np
        n = self.n_reads
        return np.ctypeslib.as_array(C.cast(self.reads_ptr, C.POINTER(C.c_uint8)), shape=(n * C.sizeof(Read),)) if n else np.zeros(0, np.uint8)

    def pool_numpy(self):
        import numpy as np
        n = self.pool_bytes
        return np.ctypeslib.as_array(C.cast(self.pool_ptr, C.POINTER(C.c_uint8)), shape=(n,)) if n else np.zeros(0, np.uint8)

    def pin(self):
        rc = lib().rvh_batch_pin(self._h)
        if rc != 0:
            raise RabbitVarError("rvh_batch_pin: " + lib().rvh_last_error().decode())

    def make_regions(self, starts, ends, ref_extension=1200, read_offset=0, n_reads_sample=-1):
        n = len(starts)
        s = (C.c_int32 * n)(*starts)
        e = (C.c_int32 * n)(*ends)
        out = (Region * n)()
        rc = lib().rvh_make_regions(self._h, s, e, n, self.chr_len, ref_extension, read_offset, n_reads_sample, out)
        if rc != 0:
            raise RabbitVarError(f"rvh_make_regions failed ({rc})")
        return out

    def close(self):
        if self._h:
            lib().rvh_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fetch_ref(fasta_path, chrom, lo, hi):
    buf = C.create_string_buffer(hi - lo + 2)
    n = lib().rvh_fetch_ref(os.fsencode(fasta_path), chrom.encode(), lo, hi, buf)
    if n < 0:
        raise RabbitVarError("rvh_fetch_ref: " + lib().rvh_last_error().decode())
    return buf.raw[:n]


class Context:
    """One GPU context (= one host thread's dataPool + BAM handles in the reference)."""

    def __init__(self, device=0, params=None, limits=None):
        L = lib()
        if L.rv_device_count() <= 0:
            raise RabbitVarError("no CUDA device visible: rabbitvar_b200 has no CPU fallback")
        self.params = params or default_params()
        self.limits = limits or default_limits()
        h = C.c_void_p()
        rc = L.rv_create(C.byref(h), device, C.byref(self.params), C.byref(self.limits))
        self._h = h
        if rc != 0:
            msg = L.rv_last_error(h).decode() if h else "rv_create failed"
            if h:
                L.rv_destroy(h)
                self._h = None
            raise RabbitVarError(f"rv_create({device}) failed ({rc}): {msg}")

    def _ck(self, rc, what):
        if rc != 0:
            raise RabbitVarError(f"{what} failed ({rc}): {lib().rv_last_error(self._h).decode()}")

    def set_params(self, params):
        self.params = params
        self._ck(lib().rv_set_params(self._h, C.byref(params)), "rv_set_params")

    def set_reference(self, ref_start, bases):
        self._ref_keep = bases
        self._ck(lib().rv_set_reference(self._h, ref_start, len(bases), C.cast(C.c_char_p(bases), C.c_void_p)),
                 "rv_set_reference")

    def push_reads(self, batch):
        rb = ReadBatch(batch.n_reads, batch.reads_ptr, batch.pool_ptr, batch.pool_bytes)
        self._ck(lib().rv_push_reads(self._h, C.byref(rb)), "rv_push_reads")

    def push_reads_ptr(self, n_reads, reads_ptr, pool_ptr, pool_bytes, device=False):
        rb = ReadBatch(n_reads, reads_ptr, pool_ptr, pool_bytes)
        f = lib().rv_push_reads_device if device else lib().rv_push_reads
        self._ck(f(self._h, C.byref(rb)), "rv_push_reads")

    def set_regions(self, regions):
        self._ck(lib().rv_set_regions(self._h, regions, len(regions)), "rv_set_regions")

    def pileup(self):
        self._ck(lib().rv_pileup(self._h), "rv_pileup")
        st = PileupStats()
        lib().rv_get_pileup_stats(self._h, C.byref(st))
        return st

    def score(self):
        self._ck(lib().rv_score(self._h), "rv_score")

    def score_positions(self, regions, positions):
        """Full records at an explicit list of (region, position) pairs (numpy int32 arrays)."""
        import numpy as np
        r = np.ascontiguousarray(regions, dtype=np.int32)
        p = np.ascontiguousarray(positions, dtype=np.int32)
        self._ck(lib().rv_score_positions(self._h, r.ctypes.data, p.ctypes.data, len(r)), "rv_score_positions")

    def sync(self):
        self._ck(lib().rv_sync(self._h), "rv_sync")

    def set_lazy(self, on):
        self._ck(lib().rv_set_lazy(self._h, 1 if on else 0), "rv_set_lazy")

    def pileup_enqueue(self):
        """rv_pileup without reading the statistics back (lazy mode keeps it free of host synchronisation)."""
        self._ck(lib().rv_pileup(self._h), "rv_pileup")

    def fetch_variants(self):
        p = C.POINTER(Variant)()
        n = C.c_int64()
        self._ck(lib().rv_fetch_variants(self._h, C.byref(p), C.byref(n)), "rv_fetch_variants")
        return p, n.value

    def fetch_events(self):
        p = C.POINTER(Event)()
        n = C.c_int64()
        self._ck(lib().rv_fetch_events(self._h, C.byref(p), C.byref(n)), "rv_fetch_events")
        return p, n.value

    def fetch_tables(self, region):
        import numpy as np
        c = C.POINTER(C.c_uint32)()
        v = C.POINTER(C.c_uint32)()
        fp, npos = C.c_int32(), C.c_int32()
        self._ck(lib().rv_fetch_tables(self._h, region, C.byref(c), C.byref(v), C.byref(fp), C.byref(npos)),
                 "rv_fetch_tables")
        counts = np.ctypeslib.as_array(c, shape=(npos.value, 4, 8)).copy()
        cov = np.ctypeslib.as_array(v, shape=(npos.value,)).copy()
        return counts, cov, fp.value

    def fisher_exact(self, tables):
        import numpy as np
        t = np.ascontiguousarray(tables, dtype=np.int32).reshape(-1, 4)
        out = np.zeros((t.shape[0], 3), dtype=np.float64)
        self._ck(lib().rv_fisher_exact(self._h, t.ctypes.data, t.shape[0], out.ctypes.data), "rv_fisher_exact")
        return out

    def kernel_ms(self):
        a, b = C.c_float(), C.c_float()
        lib().rv_last_kernel_ms(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def pileup_split_ms(self):
        """(classify, tile index + gather, walk) device milliseconds of the last pileup."""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        lib().rv_last_pileup_split_ms(self._h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def pileup_stage_ms(self):
        """(classify, tile index + gather, walk, apply) device milliseconds of the last pileup."""
        out = (C.c_float * 4)()
        lib().rv_last_pileup_stage_ms(self._h, out)
        return tuple(out)

    def timer_start(self):
        self._ck(lib().rv_timer_start(self._h), "rv_timer_start")

    def timer_stop(self):
        ms = C.c_float()
        self._ck(lib().rv_timer_stop(self._h, C.byref(ms)), "rv_timer_stop")
        return ms.value

    def apply_patch_raw(self, entries_ptr, n_entries):
        self._ck(lib().rv_apply_patch(self._h, entries_ptr, n_entries, None, None, None, 0), "rv_apply_patch")

    def launch_count(self):
        return lib().rv_launch_count(self._h)

    def n_variants(self):
        return lib().rv_variant_count(self._h)

    def install_patch_from_events(self, batch, regions, ref_bases, ref_lo):
        rc = lib().rvh_install_patch(self._h, C.byref(self.params), batch._h, regions, len(regions),
                                     C.cast(C.c_char_p(ref_bases), C.c_void_p), ref_lo, len(ref_bases))
        if rc != 0:
            raise RabbitVarError(f"rvh_install_patch failed ({rc}): {lib().rvh_last_error().decode()}")

    def call_regions_range(self, batch, regions, lo, hi, ref_bases, ref_lo, sample, chrom, push_reference=False,
                           push_reads=True):
        """call_regions over regions[lo:hi] of a ctypes Region array."""
        sub = (Region * (hi - lo)).from_address(C.addressof(regions) + lo * C.sizeof(Region))
        return self.call_regions(batch, sub, ref_bases, ref_lo, sample, chrom, push_reference, push_reads=push_reads)

    def call_regions(self, batch, regions, ref_bases, ref_lo, sample, chrom, push_reference=True, params=None,
                     push_reads=True):
        """Host buffers in, TSV text out: the batch-level replacement of one_region_run."""
        out = C.c_char_p()
        n = C.c_int64()
        tm = Timing()
        p = params or self.params
        rc = lib().rvh_call_regions(self._h, C.byref(p), batch._h, regions, len(regions),
                                    C.cast(C.c_char_p(ref_bases), C.c_void_p), ref_lo, len(ref_bases),
                                    (1 if push_reference else 0) | (2 if push_reads else 0), sample.encode(),
                                    chrom.encode(), C.byref(out),
                                    C.byref(n), C.byref(tm))
        if rc != 0:
            raise RabbitVarError(f"rvh_call_regions failed ({rc}): {lib().rvh_last_error().decode()}")
        return C.string_at(out, n.value).decode(), tm

    def close(self):
        if getattr(self, "_h", None):
            lib().rv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def inflate_block(data, out_len):
    """One raw-DEFLATE stream through the loader's own decoder (csrc/io/fast_inflate.hpp). None when it refuses."""
    buf = C.create_string_buffer(max(1, out_len))
    n = lib().rvh_inflate_block(C.cast(C.c_char_p(data), C.c_void_p), len(data), C.cast(buf, C.c_void_p), out_len)
    return None if n < 0 else buf.raw[:n]


def crc32(data):
    return lib().rvh_crc32(C.cast(C.c_char_p(data), C.c_void_p), len(data))


def run_files(params, fasta, bam, regions, bam2=None, sample="S", decode_threads=4, gpus=1, first_device=0):
    """Files in, TSV text out (rvh_run_files).  regions: [(chr, start1, end1, gene)].  Returns (rc, text, cov_info)."""
    n = len(regions)
    chr_a = (C.c_char_p * n)(*[r[0].encode() for r in regions])
    gene_a = (C.c_char_p * n)(*[(r[3] if len(r) > 3 else r[0]).encode() for r in regions])
    st = (C.c_int32 * n)(*[r[1] for r in regions])
    en = (C.c_int32 * n)(*[r[2] for r in regions])
    out, olen = C.c_void_p(), C.c_int64()
    cov = (C.c_double * 4)()
    rc = lib().rvh_run_files(C.byref(params), os.fsencode(fasta), os.fsencode(bam), os.fsencode(bam2) if bam2 else None,
                             sample.encode(), n, chr_a, st, en, gene_a, decode_threads, gpus, first_device, C.byref(out),
                             C.byref(olen), cov)
    text = C.string_at(out, olen.value).decode() if out.value else ""
    if out.value:
        lib().rvh_free(out)
    return rc, text, tuple(cov)
