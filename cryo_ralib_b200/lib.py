"""ctypes binding of libcryo_ralib.so -- the same mechanism the reference uses to
load cuda/gpu_aln_pack.so (test_mref_gpu_align.py:91-97).  There is no CPU
fallback: if the CUDA library is missing or a call fails this raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libcryo_ralib.so")


class LibraryMissing(RuntimeError):
    pass


class CraError(RuntimeError):
    pass


class CraConfig(C.Structure):
    _fields_ = [("nx", C.c_int), ("ir", C.c_int), ("ou", C.c_int), ("rs", C.c_int),
                ("max_particles", C.c_int), ("max_refs", C.c_int),
                ("max_range", C.c_float), ("step", C.c_float),
                ("normalize_ring", C.c_int), ("row_batch", C.c_int)]


class CraSearch(C.Structure):
    _fields_ = [("cx", C.c_float), ("cy", C.c_float), ("xl", C.c_float), ("xr", C.c_float),
                ("yl", C.c_float), ("yr", C.c_float)]


class CraResult(C.Structure):
    _fields_ = [("ang", C.c_float), ("sxs", C.c_float), ("sys", C.c_float), ("mirror", C.c_int),
                ("iref", C.c_int), ("peak", C.c_float), ("sx", C.c_float), ("sy", C.c_float)]


class CraAlignStats(C.Structure):
    _fields_ = [("ms_polar", C.c_float), ("ms_ccf", C.c_float), ("ms_final", C.c_float), ("ms_total", C.c_float),
                ("launches", C.c_long), ("alignments", C.c_long), ("rows", C.c_long)]


# the reference's struct mirrors (test_mref_gpu_align.py:112-131)
class AlignConfig(C.Structure):
    _fields_ = [("sbj_num", C.c_uint), ("ref_num", C.c_uint), ("img_dim", C.c_uint),
                ("ring_num", C.c_uint), ("ring_len", C.c_uint),
                ("shift_step", C.c_float), ("shift_rng_x", C.c_float), ("shift_rng_y", C.c_float)]


class AlignParam(C.Structure):
    _fields_ = [("sbj_id", C.c_int), ("ref_id", C.c_int), ("shift_x", C.c_float), ("shift_y", C.c_float),
                ("angle", C.c_float), ("mirror", C.c_bool)]


SEARCH_DTYPE = np.dtype([("cx", "f4"), ("cy", "f4"), ("xl", "f4"), ("xr", "f4"), ("yl", "f4"), ("yr", "f4")])
RESULT_DTYPE = np.dtype([("ang", "f4"), ("sxs", "f4"), ("sys", "f4"), ("mirror", "i4"), ("iref", "i4"),
                         ("peak", "f4"), ("sx", "f4"), ("sy", "f4")])

CORE_SYMBOLS = ["cra_create", "cra_destroy", "cra_last_error", "cra_ring_info", "cra_upload_particles",
                "cra_upload_particles_dev", "cra_upload_particles_async", "cra_upload_wait", "cra_mref_search_request",
                "cra_compose_result", "cra_reffree_search_request", "cra_fit_tanh", "cra_set_refs", "cra_align", "cra_align_bound", "cra_refs_from_sums", "cra_filter_refs",
                "cra_class_fsc", "cra_put_ref", "cra_filter_center_refs", "cra_prepare_refs",
                "cra_get_refs", "cra_accumulate", "cra_accumulate_d", "cra_transform_d", "cra_zero_sums",
                "cra_sums_device_ptr", "cra_get_sums", "cra_transform", "cra_transform_dev", "cra_polar_spectrum", "cra_ref_spectrum", "cra_batch_row_spectrum",
                "cra_ccf_curves", "cra_last_align_stats", "cra_set_timing", "cra_set_normalize_ring", "cra_set_step",
                "cra_row_batch", "cra_device_images_ptr", "cra_stream", "cra_measure_fp32_peak"]
LEGACY_SYMBOLS = ["print_gpu_info", "pre_align_size_check", "pre_align_init", "pre_align_fetch", "reset_shifts",
                  "mref_align_run", "mref_align_run_m", "get_num_ref", "pre_align_run", "pre_align_run_m", "gpu_clear",
                  "ref_free_alignment_2D_init", "ref_free_alignment_2D_size_check", "ref_free_alignment_2D",
                  "ref_free_alignment_2D_filter_references"]

_LIB = None


def load_library(path=None):
    """Load the C-ABI library; raises LibraryMissing (never falls back)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or os.environ.get("CRA_LIBRARY") or SO_PATH        # CRA_LIBRARY: another build of the same library (experiments)
    if not os.path.exists(path):
        raise LibraryMissing("%s not found: build it with `python -m cryo_ralib_b200.build` "
                             "(there is no CPU fallback)" % path)
    L = C.CDLL(path)
    vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
    L.cra_create.argtypes = [C.POINTER(CraConfig), C.c_int, C.POINTER(vp)]
    L.cra_destroy.argtypes = [vp]
    L.cra_last_error.restype = C.c_char_p
    L.cra_ring_info.argtypes = [vp, ip, ip, ip, vp]
    L.cra_upload_particles.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    L.cra_upload_particles_dev.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    L.cra_upload_particles_async.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    L.cra_upload_wait.argtypes = [vp]
    L.cra_mref_search_request.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, vp]
    L.cra_compose_result.argtypes = [C.c_int, vp, vp, vp, vp]
    L.cra_reffree_search_request.argtypes = [C.c_int, vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                             vp, vp, vp]
    L.cra_fit_tanh.argtypes = [C.c_int, vp, vp, C.c_double, C.c_double, C.c_double, C.c_double, vp]
    L.cra_set_refs.argtypes = [vp, vp, C.c_int, C.c_int]
    L.cra_align.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.cra_align_bound.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    L.cra_refs_from_sums.argtypes = [vp, C.c_int]
    L.cra_filter_refs.argtypes = [vp, C.c_float, C.c_float, C.c_int]
    L.cra_get_refs.argtypes = [vp, vp]
    L.cra_class_fsc.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_int), vp, vp, vp, vp]
    L.cra_put_ref.argtypes = [vp, C.c_int, vp]
    L.cra_filter_center_refs.argtypes = [vp, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_int, vp]
    L.cra_prepare_refs.argtypes = [vp, C.c_int]
    L.cra_accumulate.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_long]
    L.cra_accumulate_d.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_long]
    L.cra_zero_sums.argtypes = [vp]
    L.cra_sums_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.cra_get_sums.argtypes = [vp, vp, vp]
    L.cra_transform.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.cra_transform_d.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.cra_transform_dev.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.cra_polar_spectrum.argtypes = [vp, C.c_int, C.c_float, C.c_float, vp]
    L.cra_ref_spectrum.argtypes = [vp, C.c_int, vp]
    L.cra_batch_row_spectrum.argtypes = [vp, C.c_int, vp, ip]
    L.cra_ccf_curves.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.c_int, vp, vp]
    L.cra_last_align_stats.argtypes = [vp, C.POINTER(CraAlignStats)]
    L.cra_set_timing.argtypes = [vp, C.c_int]
    L.cra_set_normalize_ring.argtypes = [vp, C.c_int]
    L.cra_set_step.argtypes = [vp, C.c_float]
    L.cra_row_batch.argtypes = [vp]
    L.cra_device_images_ptr.argtypes = [vp, C.POINTER(vp)]
    L.cra_stream.argtypes = [vp]
    L.cra_stream.restype = vp
    L.cra_measure_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    # legacy layer, restypes as the reference drivers set them (test_mref_gpu_align.py:95-97)
    L.pre_align_init.restype = C.c_ulonglong
    L.pre_align_init.argtypes = [C.c_uint, C.POINTER(AlignConfig), C.c_uint]
    L.pre_align_size_check.restype = C.c_bool
    L.pre_align_size_check.argtypes = [C.c_uint, C.POINTER(AlignConfig), C.c_uint, C.c_float, C.c_bool]
    L.pre_align_fetch.argtypes = [C.POINTER(C.POINTER(C.c_float)), C.c_uint, C.c_char_p]
    L.reset_shifts.argtypes = [C.c_float, C.c_float]
    L.mref_align_run.restype = C.c_ulonglong
    L.mref_align_run.argtypes = [C.c_int, C.c_int]
    L.mref_align_run_m.restype = C.POINTER(C.c_float)
    L.mref_align_run_m.argtypes = [C.c_int, C.c_int]
    L.get_num_ref.restype = C.POINTER(C.c_int)
    L.pre_align_run.argtypes = [C.c_int, C.c_int]
    L.pre_align_run_m.restype = C.c_ulonglong
    L.pre_align_run_m.argtypes = [C.c_int, C.c_int]
    # gpu_isac's class-bound variant (gpu_aln_noref.h:94-109)
    L.ref_free_alignment_2D_init.restype = C.c_ulonglong
    L.ref_free_alignment_2D_init.argtypes = [C.POINTER(AlignConfig), C.POINTER(C.POINTER(C.c_float)),
                                             C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int), C.c_uint]
    L.ref_free_alignment_2D_size_check.restype = C.c_bool
    L.ref_free_alignment_2D_size_check.argtypes = [C.POINTER(AlignConfig), C.c_uint, C.c_float, C.c_bool]
    L.ref_free_alignment_2D.argtypes = []
    L.ref_free_alignment_2D_filter_references.argtypes = [C.c_float, C.c_float]
    if path == SO_PATH or path == os.environ.get("CRA_LIBRARY"):
        _LIB = L
    return L


class Engine(object):
    """Thin object wrapper over the cra_* core (one context = one GPU)."""

    def __init__(self, nx, ou, xr, yr=None, ts=1.0, ir=1, rs=1, max_particles=1, max_refs=1,
                 normalize_ring=True, device=0, row_batch=0):
        self.L = load_library()
        yr = xr if yr is None else yr
        self.cfg = CraConfig(int(nx), int(ir), int(ou), int(rs), int(max_particles), int(max_refs),
                             float(max(xr, yr)), float(ts), int(bool(normalize_ring)), int(row_batch))
        self.h = C.c_void_p()
        self._ck(self.L.cra_create(C.byref(self.cfg), int(device), C.byref(self.h)))
        self.nx, self.ou, self.device = int(nx), int(ou), int(device)
        self.max_particles, self.max_refs = int(max_particles), int(max_refs)
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.cra_ring_info(self.h, C.byref(a), C.byref(b), C.byref(c), None))
        self.nring, self.lcirc, self.maxrin = a.value, b.value, c.value
        numr = np.zeros(3 * self.nring, np.int32)
        self._ck(self.L.cra_ring_info(self.h, None, None, None, numr.ctypes.data))
        self.numr = numr
        self.R = 0

    def _ck(self, rc):
        if rc != 0:
            raise CraError(self.L.cra_last_error().decode())

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.cra_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data
    def upload_particles(self, images, first=0, subtract_mask_mean=True):
        images = np.ascontiguousarray(images, np.float32)
        assert images.ndim == 3 and images.shape[1] == self.nx and images.shape[2] == self.nx
        self._ck(self.L.cra_upload_particles(self.h, images.ctypes.data, int(first), images.shape[0],
                                             int(subtract_mask_mean)))

    def upload_particles_ptr(self, host_ptr, n, first=0, subtract_mask_mean=True):
        self._ck(self.L.cra_upload_particles(self.h, host_ptr, int(first), int(n), int(subtract_mask_mean)))

    def upload_particles_async(self, host_ptr, n, first=0, subtract_mask_mean=True):
        """Queue a pinned-host upload on the copy stream; align()/accumulate() of an overlapping particle
        range wait for it on the device.  The buffer must stay valid until upload_wait()."""
        self._ck(self.L.cra_upload_particles_async(self.h, host_ptr, int(first), int(n), int(subtract_mask_mean)))

    def upload_wait(self):
        self._ck(self.L.cra_upload_wait(self.h))

    def upload_particles_dev(self, dev_ptr, n, first=0, subtract_mask_mean=True):
        self._ck(self.L.cra_upload_particles_dev(self.h, dev_ptr, int(first), int(n), int(subtract_mask_mean)))

    def set_refs(self, refs, normalize_mask=True):
        refs = np.ascontiguousarray(refs, np.float32)
        self._ck(self.L.cra_set_refs(self.h, refs.ctypes.data, refs.shape[0], int(normalize_mask)))
        self.R = refs.shape[0]

    # ---- hot path
    def align(self, start, stop, search):
        """search: structured array (SEARCH_DTYPE) of length stop-start.  Returns RESULT_DTYPE array."""
        search = np.ascontiguousarray(search, SEARCH_DTYPE)
        assert search.shape[0] == stop - start
        out = np.zeros(stop - start, RESULT_DTYPE)
        self._ck(self.L.cra_align(self.h, int(start), int(stop), search.ctypes.data, out.ctypes.data))
        return out

    def align_bound(self, start, stop, search, class_of):
        """Class-bound alignment (gpu_isac ref_free_alignment_2D): particle p against reference class_of[p] only."""
        search = np.ascontiguousarray(search, SEARCH_DTYPE)
        class_of = np.ascontiguousarray(class_of, np.int32)
        assert search.shape[0] == stop - start and class_of.shape[0] == stop - start
        out = np.zeros(stop - start, RESULT_DTYPE)
        self._ck(self.L.cra_align_bound(self.h, int(start), int(stop), search.ctypes.data, class_of.ctypes.data,
                                        out.ctypes.data))
        return out

    def refs_from_sums(self, normalize_mask=False):
        """References <- class averages of the accumulated sums, on the device."""
        self._ck(self.L.cra_refs_from_sums(self.h, int(normalize_mask)))

    def filter_refs(self, cutoff, falloff, normalize_mask=False):
        """Tangent low-pass (filt_tanl) of the current references, on the device."""
        self._ck(self.L.cra_filter_refs(self.h, float(cutoff), float(falloff), int(normalize_mask)))

    def class_fsc(self, masked=False, min_members=4, write_avg=True, avg_div=0.0):
        """Device half of the reference update: class averages into the reference slots and fsc(even, odd) per class
        (test_mref.py:252-256).  Returns (freq [nsh], fsc [R][nsh], n [nsh], counts [R])."""
        nsh = C.c_int()
        self._ck(self.L.cra_class_fsc(self.h, int(masked), int(min_members), int(write_avg), float(avg_div), C.byref(nsh),
                                      None, None, None, None))
        freq = np.zeros(nsh.value); n = np.zeros(nsh.value)
        fsc = np.zeros((self.R, nsh.value)); counts = np.zeros(self.R, np.float32)
        self._ck(self.L.cra_class_fsc(self.h, int(masked), int(min_members), int(write_avg), float(avg_div), C.byref(nsh),
                                      freq.ctypes.data, fsc.ctypes.data, n.ctypes.data, counts.ctypes.data))
        return freq, fsc, n, counts

    def put_ref(self, iref, img):
        img = np.ascontiguousarray(img, np.float32)
        assert img.shape == (self.nx, self.nx)
        self._ck(self.L.cra_put_ref(self.h, int(iref), img.ctypes.data))

    def filter_center_refs(self, cutoff, falloff, mode=1, shift=(0.0, 0.0), normalize_mask=True):
        """filt_tanl + centring (mode 1: phase centre of gravity; 2: the given shift; 0: none) + normalize.mask of every
        reference on the device (ref_ali2d, test_mref.py:273-284).  Returns the shifts removed, [R][2]."""
        cs = np.zeros((self.R, 2), np.float32)
        self._ck(self.L.cra_filter_center_refs(self.h, float(cutoff), float(falloff), int(mode), float(shift[0]), float(shift[1]),
                                               int(normalize_mask), cs.ctypes.data))
        return cs

    def prepare_refs(self, normalize_mask=True):
        """Reference preparation (test_mref.py:170-175) of the references already on the device."""
        self._ck(self.L.cra_prepare_refs(self.h, int(normalize_mask)))

    def update_refs_device(self, center=1, reseed=None, fetch=True):
        """The whole mref reference update (test_mref.py:238-286) with the class sums staying on the GPU: the host sees
        only R FSC curves of nx/2+1 numbers and fits the tangent filter (refupdate.fit_tanh -> cra_fit_tanh).
        Returns (refs or None, info) like refupdate.update_refs."""
        from . import refupdate as ru
        freq, fsc, n, counts = self.class_fsc(masked=False, min_members=4, write_avg=True)
        alive = counts >= 4
        if not alive.any():
            raise RuntimeError("every reference vanished (all classes have < 4 members)")
        reseeded = [int(j) for j in np.nonzero(~alive)[0]]
        for j in reseeded:
            self.put_ref(j, reseed(j))
        keep = n > 0
        last = int(np.nonzero(alive)[0][-1])
        acc = fsc[alive].sum(axis=0)
        # the reference hands ref_ali2d the LAST surviving class's fsc lists with the class-averaged curve written into
        # them (test_mref.py:258-271); the curve stays that class's own when the averaged one sums to zero
        curve = acc / float(alive.sum()) if acc.sum() != 0 else fsc[last]
        frsc = [list(freq[keep]), list(curve[keep]), list(n[keep])]
        fl, aa = ru.fit_tanh(frsc)
        aa = min(aa, 0.2)
        fl = max(min(0.4, fl), 0.12)
        if center not in (0, 1):
            raise NotImplementedError("center methods other than 0/1 are off the path")
        cs = self.filter_center_refs(fl, aa, mode=int(center), normalize_mask=True)
        info = dict(frsc=frsc, reseeded=reseeded, cs=[list(map(float, c)) for c in cs], filter=(fl, aa),
                    class_fsc={int(j): [list(freq[keep]), list(fsc[j][keep]), list(n[keep])] for j in np.nonzero(alive)[0]})
        return (self.get_refs() if fetch else None), info

    def get_refs(self):
        out = np.zeros((self.R, self.nx, self.nx), np.float32)
        self._ck(self.L.cra_get_refs(self.h, out.ctypes.data))
        return out

    def accumulate(self, start, stop, params, iref, global_offset=0):
        params = np.ascontiguousarray(params, np.float64)       # doubles, as the reference's Python holds them
        iref = np.ascontiguousarray(iref, np.int32)
        assert params.shape == (stop - start, 4) and iref.shape[0] == stop - start
        self._ck(self.L.cra_accumulate_d(self.h, int(start), int(stop), params.ctypes.data, iref.ctypes.data,
                                         int(global_offset)))

    def zero_sums(self):
        self._ck(self.L.cra_zero_sums(self.h))

    def get_sums(self):
        R = self.max_refs
        sums = np.zeros((R, 2, self.nx, self.nx), np.float32)
        counts = np.zeros(R, np.float32)
        self._ck(self.L.cra_get_sums(self.h, sums.ctypes.data, counts.ctypes.data))
        return sums, counts

    def get_sums_counts(self):
        """The class sizes only (the [R] tail of the sums buffer)."""
        counts = np.zeros(self.max_refs, np.float32)
        self._ck(self.L.cra_get_sums(self.h, None, counts.ctypes.data))
        return counts

    def sums_device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.cra_sums_device_ptr(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def transform(self, start, stop, params):
        params = np.ascontiguousarray(params, np.float64)
        out = np.zeros((stop - start, self.nx, self.nx), np.float32)
        self._ck(self.L.cra_transform_d(self.h, int(start), int(stop), params.ctypes.data, out.ctypes.data))
        return out

    # ---- stage-level (tests)
    def polar_spectrum(self, particle, cx, cy):
        out = np.zeros(self.lcirc, np.float32)
        self._ck(self.L.cra_polar_spectrum(self.h, int(particle), cx, cy, out.ctypes.data))
        return out

    def batch_row_spectrum(self, row):
        """Spectrum of one row of the last batch align() processed, and which row kernel wrote it."""
        out = np.zeros(self.lcirc, np.float32)
        k = C.c_int()
        self._ck(self.L.cra_batch_row_spectrum(self.h, int(row), out.ctypes.data, C.byref(k)))
        return out, k.value

    def ref_spectrum(self, iref):
        out = np.zeros(self.lcirc, np.float32)
        self._ck(self.L.cra_ref_spectrum(self.h, int(iref), out.ctypes.data))
        return out

    def ccf_curves(self, particle, cx, cy, iref):
        q = np.zeros(self.maxrin, np.float32)
        t = np.zeros(self.maxrin, np.float32)
        self._ck(self.L.cra_ccf_curves(self.h, int(particle), cx, cy, int(iref), q.ctypes.data, t.ctypes.data))
        return q, t

    # ---- knobs / stats
    def set_timing(self, on=True):
        self._ck(self.L.cra_set_timing(self.h, int(on)))

    def set_normalize_ring(self, on):
        self._ck(self.L.cra_set_normalize_ring(self.h, int(on)))

    def set_step(self, ts):
        self._ck(self.L.cra_set_step(self.h, float(ts)))

    def stats(self):
        s = CraAlignStats()
        self._ck(self.L.cra_last_align_stats(self.h, C.byref(s)))
        return dict(ms_polar=s.ms_polar, ms_ccf=s.ms_ccf, ms_final=s.ms_final, ms_total=s.ms_total,
                    launches=s.launches, alignments=s.alignments, rows=s.rows)

    def measure_fp32_peak(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.cra_measure_fp32_peak(self.device, C.byref(a), C.byref(b)))
        return a.value, b.value
