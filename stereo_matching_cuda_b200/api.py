"""ctypes binding of include/stereo_b200.h and a numpy/torch-friendly mirror of the
reference's host stage functions (same names, argument meaning and in/out conventions;
paths below are relative to stereo_matching_cuda/ in the reference).

Host entry points take numpy arrays (HOST memory), block, and return numpy arrays -- the
reference's calling convention.  `*_dev` methods take torch CUDA tensors (or raw device
pointers) and run asynchronously on the context's stream; torch is only used for device
memory and streams.  Every failure raises StereoB200Error with the library's message; a
missing or unloadable libstereo_b200.so raises immediately (there is no fallback).
"""
import ctypes as C
import os

import numpy as np

GUIDE_GRAY, GUIDE_RGB = 0, 1
BOX_SLIDING, BOX_SAT = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))


class StereoB200Error(RuntimeError):
    pass


class Params(C.Structure):
    """sb200_params: the reference's SystemIncludes.h:6-24 macros as runtime data."""

    _fields_ = [
        ("dmin", C.c_int),
        ("dmax", C.c_int),
        ("radius", C.c_int),
        ("d_lr", C.c_int),
        ("eps", C.c_double),
        ("alpha", C.c_float),
        ("th_color", C.c_float),
        ("th_grad", C.c_float),
        ("r_w", C.c_double),
        ("g_w", C.c_double),
        ("b_w", C.c_double),
        ("guide_mode", C.c_int),
        ("box_mode", C.c_int),
    ]

    @property
    def size_d(self):
        return self.dmax - self.dmin + 1


class WMedianParams(C.Structure):
    """sb200_wmedian_params"""
    _fields_ = [("radius", C.c_int), ("sigma_space", C.c_float), ("sigma_color", C.c_float)]


class _Outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("disp_left", "disp_right", "occlusion", "filled", "best_left", "best_right",
                                          "gray_left", "gray_right", "mean_left", "mean_right", "subpixel_left")]


class _LabelsI16(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("disp_left", "disp_right", "occlusion", "filled")]


class _Strip(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("y0", "rows", "halo_top", "halo_bot", "frame_h")]


def lib_path():
    # SB200_LIB lets a developer A/B another build of the same ABI; the default is the in-tree library
    return os.environ.get("SB200_LIB") or os.path.join(_HERE, "libstereo_b200.so")


_LIB = None


def load_library(path=None):
    """Loads libstereo_b200.so (built by `python -m stereo_matching_cuda_b200.build`)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or lib_path()
    if not os.path.exists(path):
        raise StereoB200Error(f"{path} is missing: run `python -m stereo_matching_cuda_b200.build` (nvcc, sm_100a)")
    L = C.CDLL(path)
    vp, ip, fp = C.c_void_p, C.c_int, C.c_float
    PP = C.POINTER(Params)
    sig = {
        "sb200_default_params": (None, [PP]),
        "sb200_ctx_create": (ip, [ip, C.POINTER(vp)]),
        "sb200_ctx_destroy": (None, [vp]),
        "sb200_ctx_set_stream": (ip, [vp, vp]),
        "sb200_ctx_synchronize": (ip, [vp]),
        "sb200_last_error": (C.c_char_p, [vp]),
        "sb200_launch_count": (C.c_uint64, [vp]),
        "sb200_ctx_rgb_kernel": (ip, [vp]),
        "sb200_pipeline_path": (ip, [vp, PP]),
        "sb200_ctx_gray_kernel": (ip, [vp]),
        "sb200_ctx_set_gray_kernel": (ip, [vp, ip]),
        "sb200_version": (C.c_char_p, []),
        "sb200_rgb_to_grayscale": (ip, [vp, PP, vp, ip, ip, vp]),
        "sb200_compute_cost": (ip, [vp, PP, vp, vp, vp, ip, ip, ip, ip, ip]),
        "sb200_x_derivative": (ip, [vp, vp, vp, ip, ip]),
        "sb200_integral": (ip, [vp, vp, vp, ip, ip]),
        "sb200_box_filter_sat": (ip, [vp, PP, vp, vp, ip, ip]),
        "sb200_box_filter": (ip, [vp, PP, vp, vp, ip, ip]),
        "sb200_filter": (ip, [vp, PP, vp, ip, ip, vp, vp]),
        "sb200_compute_guided_filter": (ip, [vp, PP, vp, vp, vp, vp, vp, ip, ip, ip, ip]),
        "sb200_winner_take_all": (ip, [vp, vp, vp, vp, ip, ip]),
        "sb200_detect_occlusion": (ip, [vp, PP, vp, vp, ip, vp, vp, ip, ip]),
        "sb200_fill_occlusion": (ip, [vp, vp, ip, ip, fp]),
        "sb200_rgb_to_grayscale_dev": (ip, [vp, PP, vp, ip, ip, vp]),
        "sb200_compute_cost_dev": (ip, [vp, PP, vp, vp, vp, ip, ip, ip]),
        "sb200_integral_dev": (ip, [vp, vp, vp, ip, ip]),
        "sb200_box_filter_dev": (ip, [vp, PP, vp, vp, ip, ip]),
        "sb200_compute_guided_filter_dev": (ip, [vp, PP, vp, vp, vp, vp, vp, ip, ip, ip, ip]),
        "sb200_winner_take_all_dev": (ip, [vp, vp, vp, vp, ip, ip]),
        "sb200_detect_occlusion_dev": (ip, [vp, PP, vp, vp, ip, ip, ip]),
        "sb200_fill_occlusion_dev": (ip, [vp, vp, ip, ip, fp]),
        "sb200_default_wmedian_params": (None, [vp]),
        "sb200_weighted_median": (ip, [vp, vp, vp, vp, vp, vp, vp, ip, ip]),
        "sb200_weighted_median_dev": (ip, [vp, vp, vp, vp, vp, vp, vp, ip, ip]),
        "sb200_pipeline_dev": (ip, [vp, PP, vp, vp, ip, ip, ip, C.POINTER(_Outputs)]),
        "sb200_pipeline": (ip, [vp, PP, vp, vp, ip, ip, ip, C.POINTER(_Outputs)]),
        "sb200_last_exchange_ms": (ip, [vp, C.POINTER(C.c_float)]),
        "sb200_write_mat": (ip, [vp, vp, vp, ip, ip]),
        "sb200_write_mat_dev": (ip, [vp, vp, vp, ip, ip]),
        "sb200_fl_to_ch2_dev": (ip, [vp, vp, vp, ip, ip, ip]),
        "sb200_pipeline_strips_nccl": (ip, [vp, PP, vp, ip, ip, vp, vp, ip, ip, ip, ip, ip, C.POINTER(_Outputs)]),
        "sb200_strip_rows": (ip, [ip, ip, ip, C.POINTER(ip), C.POINTER(ip)]),
        "sb200_pipeline_batch": (ip, [vp, PP, vp, vp, ip, ip, ip, ip, C.POINTER(_Outputs)]),
        "sb200_pipeline_batch_dev": (ip, [vp, PP, vp, vp, ip, ip, ip, ip, C.POINTER(_Outputs)]),
        "sb200_pipeline_strip_dev": (ip, [vp, PP, vp, vp, ip, ip, C.POINTER(_Strip), C.POINTER(_Outputs)]),
        "sb200_strip_halo_rows": (ip, [PP]),
        "sb200_view_disparity_dev": (ip, [vp, PP, vp, vp, ip, ip, ip, ip, vp, vp, vp]),
        "sb200_pipeline_batch_i16": (ip, [vp, PP, vp, vp, ip, ip, ip, ip, vp, vp]),
        "sb200_view_volume_dev": (ip, [vp, PP, vp, vp, ip, ip, ip, ip, vp, vp, vp]),
        "sb200_subpixel_refine_dev": (ip, [vp, vp, vp, vp, vp, vp, ip, ip, ip, ip]),
        "sb200_lr_check_fill_dev": (ip, [vp, PP, vp, vp, ip, ip, ip, fp, vp, vp]),
        "sb200_ctx_enable_timing": (ip, [vp, ip]),
        "sb200_last_timing": (ip, [vp] + [C.POINTER(fp)] * 4),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here means the .so does not export the header's symbol
        fn.restype, fn.argtypes = res, args
    L._sb_signatures = sig
    if path == lib_path():
        _LIB = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily by tests from the header


def default_params(**kw):
    p = Params()
    load_library().sb200_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown parameter {k}")
        setattr(p, k, v)
    return p


def _ptr(x):
    """device/host pointer of a numpy array, torch tensor, int or None"""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    return C.c_void_p(x.data_ptr())  # torch tensor


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Context:
    """One context per host thread / GPU (main.cu:44-48 replaced).  Owns the workspace arena."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.sb200_ctx_create(device, C.byref(h))
        if rc != 0:
            raise StereoB200Error(f"sb200_ctx_create({device}) failed [{rc}]: " +
                                  self.lib.sb200_last_error(None).decode())
        self.h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "h", None):
            self.lib.sb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise StereoB200Error(f"[{rc}] " + self.lib.sb200_last_error(self.h).decode())

    def set_stream(self, stream):
        """stream: a raw cudaStream_t handle (int) or a torch.cuda.Stream"""
        handle = stream if isinstance(stream, int) else stream.cuda_stream
        self._ck(self.lib.sb200_ctx_set_stream(self.h, C.c_void_p(handle)))

    def synchronize(self):
        self._ck(self.lib.sb200_ctx_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.lib.sb200_launch_count(self.h))

    def pipeline_path(self, params=None):
        """0 = staged kernels, 1 = warp-shuffle fused kernel, 2 = tensor-core fused kernel (sb200_pipeline_path)"""
        p = params or default_params()
        rc = int(self.lib.sb200_pipeline_path(self.h, C.byref(p)))
        if rc < 0:
            raise StereoB200Error(f"no pipeline path for these parameters (code {rc})")
        return rc

    @property
    def rgb_kernel(self):
        """2: k_fused_cvf_rgb, 3: k_fused_cvf_rgb3 (SB200_RGB_KERNEL overrides the library default)"""
        return int(self.lib.sb200_ctx_rgb_kernel(self.h))

    @property
    def gray_kernel(self):
        """1: k_fused_mma (tensor-core window sums), 0: k_fused_cvf (warp-shuffle window sums)"""
        return int(self.lib.sb200_ctx_gray_kernel(self.h))

    @gray_kernel.setter
    def gray_kernel(self, which):
        self._ck(self.lib.sb200_ctx_set_gray_kernel(self.h, int(which)))

    def enable_timing(self, on=True):
        self._ck(self.lib.sb200_ctx_enable_timing(self.h, int(on)))

    def last_timing(self):
        v = [C.c_float() for _ in range(4)]
        self._ck(self.lib.sb200_last_timing(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("prep_ms", "fused_ms", "merge_ms", "occl_ms"), (x.value for x in v)))

    def last_exchange_ms(self):
        v = C.c_float()
        self._ck(self.lib.sb200_last_exchange_ms(self.h, C.byref(v)))
        return v.value

    # ---- reference stage functions, host arrays ------------------------------------------
    def rgb_to_grayscale(self, h_rgb, params=None):
        """rgb_to_grayscale.cuh:7 -- (h, w, ch>=3) uint8 -> (h, w) uint8"""
        p = params or default_params()
        rgb = _np(h_rgb, np.uint8)
        h, w, ch = rgb.shape
        gray = np.empty((h, w), np.uint8)
        self._ck(self.lib.sb200_rgb_to_grayscale(self.h, C.byref(p), _ptr(rgb), h * w, ch, _ptr(gray)))
        return gray

    def x_derivative(self, img):
        """costVolume.cuh:16 (x_derivativeOnGPU)"""
        img = _np(img, np.uint8)
        h, w = img.shape
        out = np.empty((h, w), np.float32)
        self._ck(self.lib.sb200_x_derivative(self.h, _ptr(img), _ptr(out), w, h))
        return out

    def compute_cost(self, i1, i2, dmin, params=None):
        """costVolume.cuh:7 -- returns the planar (size_d, h, w) float32 volume"""
        p = params or default_params()
        i1, i2 = _np(i1, np.uint8), _np(i2, np.uint8)
        h, w = i1.shape
        cost = np.empty((p.size_d, h, w), np.float32)
        self._ck(self.lib.sb200_compute_cost(self.h, C.byref(p), _ptr(i1), _ptr(i2), _ptr(cost), w, i2.shape[1], h,
                                             i2.shape[0], dmin))
        return cost

    def integral(self, image):
        """integral.cuh:3"""
        image = _np(image, np.float32)
        h, w = image.shape
        out = np.empty((h, w), np.float32)
        self._ck(self.lib.sb200_integral(self.h, _ptr(image), _ptr(out), w, h))
        return out

    def box_filter_sat(self, integral, params=None):
        """guidedFilter.cuh:23 (computeBoxFilterOnGPU)"""
        p = params or default_params()
        integral = _np(integral, np.float32)
        h, w = integral.shape
        out = np.empty((h, w), np.float32)
        self._ck(self.lib.sb200_box_filter_sat(self.h, C.byref(p), _ptr(integral), _ptr(out), w, h))
        return out

    def box_filter(self, image, params=None):
        p = params or default_params()
        image = _np(image, np.float32)
        h, w = image.shape
        out = np.empty((h, w), np.float32)
        self._ck(self.lib.sb200_box_filter(self.h, C.byref(p), _ptr(image), _ptr(out), w, h))
        return out

    def filter(self, image, params=None):
        """filter.cuh:12 -- returns (mean uint8, var float32) of the guide"""
        p = params or default_params()
        image = _np(image, np.uint8)
        h, w = image.shape
        mean, var = np.empty((h, w), np.uint8), np.empty((h, w), np.float32)
        self._ck(self.lib.sb200_filter(self.h, C.byref(p), _ptr(image), w, h, _ptr(mean), _ptr(var)))
        return mean, var

    def compute_guided_filter(self, i, cost, filter_cost, disp_map, dmin, params=None):
        """guidedFilter.cuh:7 -- filter_cost and disp_map are updated IN PLACE (main.cu:112-120
        initialises them); returns the uint8 mean image."""
        p = params or default_params()
        i = _np(i, np.uint8)
        cost = _np(cost, np.float32)
        if cost.ndim != 3:
            raise StereoB200Error(f"cost {cost.shape}: expected the planar (size_d, h, w) volume")
        size_d, h, w = cost.shape
        for name, a in (("filter_cost", filter_cost), ("disp_map", disp_map)):
            if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags.c_contiguous or a.shape != (h, w):
                raise StereoB200Error(f"{name}: expected a C-contiguous float32 array of shape {(h, w)} (updated in place)")
        if i.shape != (h, w):
            raise StereoB200Error(f"guide image {i.shape} does not match the cost volume's {(h, w)}")
        mean = np.empty((h, w), np.uint8)
        self._ck(self.lib.sb200_compute_guided_filter(self.h, C.byref(p), _ptr(i), _ptr(cost), _ptr(filter_cost),
                                                      _ptr(disp_map), _ptr(mean), w, h, size_d, dmin))
        return mean

    def winner_take_all(self, q, filter_cost, dmap, label):
        """guidedFilter.cuh:8 (dispSelectOnGPU) -- in place on filter_cost / dmap"""
        q = _np(q, np.float32)
        assert filter_cost.dtype == np.float32 and dmap.dtype == np.float32
        self._ck(self.lib.sb200_winner_take_all(self.h, _ptr(q), _ptr(filter_cost), _ptr(dmap), q.size, label))

    def detect_occlusion(self, disparity_left, disparity_right, d_occlusion, params=None):
        """occlusion.cuh:8 -- in place on disparity_left"""
        p = params or default_params()
        assert disparity_left.dtype == np.float32 and disparity_left.flags.c_contiguous
        dr = _np(disparity_right, np.float32)
        h, w = disparity_left.shape
        self._ck(self.lib.sb200_detect_occlusion(self.h, C.byref(p), _ptr(disparity_left), _ptr(dr), d_occlusion, None,
                                                 None, w, h))

    def fill_occlusion(self, disparity, v_min):
        """occlusion.cuh:14 -- in place"""
        assert disparity.dtype == np.float32 and disparity.flags.c_contiguous
        h, w = disparity.shape
        self._ck(self.lib.sb200_fill_occlusion(self.h, _ptr(disparity), w, h, float(v_min)))

    def weighted_median(self, gray, occlusion, filled, params=None, radius=19, sigma_space=9.0, sigma_color=25.5):
        """Beyond the reference: weighted median of the pixels the L/R check marked; see stereo_b200.h"""
        p = params or default_params()
        g = _np(gray, np.uint8)
        oc, fl = _np(occlusion, np.float32), _np(filled, np.float32)
        if g.ndim != 2 or oc.shape != g.shape or fl.shape != g.shape:
            raise StereoB200Error("weighted_median: gray, occlusion and filled must be 2-D arrays of one shape")
        wp = WMedianParams(int(radius), float(sigma_space), float(sigma_color))
        out = np.empty(g.shape, np.float32)
        h, w = g.shape
        self._ck(self.lib.sb200_weighted_median(self.h, C.byref(p), C.byref(wp), _ptr(g), _ptr(oc), _ptr(fl), _ptr(out), w, h))
        return out

    # ---- fused pipeline -----------------------------------------------------------------
    _F32 = ("disp_left", "disp_right", "occlusion", "filled", "best_left", "best_right")
    _F32_OPT = ("subpixel_left",)  # beyond the reference: only when asked for by name
    _U8 = ("gray_left", "gray_right", "mean_left", "mean_right")

    def write_mat(self, mat):
        """main.cu:13-35: the reference's min/max normalisation of a float map to 8 bits, on the GPU"""
        mat = _np(mat, np.float32)
        if mat.ndim != 2:
            raise StereoB200Error(f"write_mat: expected a 2-D float map, got {mat.shape}")
        out = np.empty(mat.shape, np.uint8)
        self._ck(self.lib.sb200_write_mat(self.h, _ptr(mat), _ptr(out), mat.shape[1], mat.shape[0]))
        return out

    def write_mat_dev(self, d_mat, d_out, w, h):
        self._dev_check(d_mat, "mat", w * h, "float32")
        self._dev_check(d_out, "out", w * h, "uint8")
        self._ck(self.lib.sb200_write_mat_dev(self.h, _ptr(d_mat), _ptr(d_out), w, h))

    def fl_to_ch2_dev(self, d_image, d_result, vmin, vmax, n):
        """occlusion.cu:230-237 (flToCh2OnGPU)"""
        self._dev_check(d_image, "image", n, "float32")
        self._dev_check(d_result, "result", n, "uint8")
        self._ck(self.lib.sb200_fl_to_ch2_dev(self.h, _ptr(d_image), _ptr(d_result), vmin, vmax, n))

    def pipeline(self, left, right, params=None, want=None):
        """main.cu:65-155 as one call on HOST arrays: (h,w[,ch]) uint8 -> dict of numpy arrays"""
        p = params or default_params()
        left, right = _np(left, np.uint8), _np(right, np.uint8)
        if left.shape != right.shape:
            raise StereoB200Error(f"left {left.shape} and right {right.shape} must have the same shape")
        h, w = left.shape[:2]
        ch = 1 if left.ndim == 2 else left.shape[2]
        if want is None:
            want = self._F32 + self._U8
            if p.guide_mode == GUIDE_RGB:  # the uint8 mean images are gray-guide debug outputs
                want = tuple(k for k in want if not k.startswith("mean_"))
        res, o = {}, _Outputs()
        for k in want:
            if k not in self._F32 + self._F32_OPT + self._U8:
                raise TypeError(f"unknown output {k}")
            res[k] = np.empty((h, w), np.uint8 if k in self._U8 else np.float32)
            setattr(o, k, res[k].ctypes.data)
        self._ck(self.lib.sb200_pipeline(self.h, C.byref(p), _ptr(left), _ptr(right), ch, w, h, C.byref(o)))
        return res

    _LABELS = ("disp_left", "disp_right", "occlusion", "filled")

    def pipeline_batch(self, lefts, rights, params=None, want=("disp_left", "disp_right", "occlusion", "filled"), out=None,
                       labels_i16=False):
        """n pairs on HOST arrays (n,h,w[,ch]) uint8 through sb200_pipeline_batch: uploads, kernels and downloads of
        consecutive pairs overlap (page-locked arrays, e.g. torch pin_memory().numpy(), are needed for the overlap).
        `out`: optional dict of preallocated (n,h,w) arrays (float32 / uint8) to fill instead of fresh ones.
        labels_i16: the four label maps come back as int16 (sb200_pipeline_batch_i16: half the device->host bytes)."""
        p = params or default_params()
        lefts, rights = _np(lefts, np.uint8), _np(rights, np.uint8)
        if lefts.shape != rights.shape or lefts.ndim not in (3, 4):
            raise StereoB200Error(f"lefts {lefts.shape} / rights {rights.shape}: expected two (n,h,w[,ch]) arrays of one shape")
        n, h, w = lefts.shape[:3]
        ch = 1 if lefts.ndim == 3 else lefts.shape[3]
        res, o, li = {}, _Outputs(), _LabelsI16()
        for k in want:
            dt = np.float32 if k in self._F32 else np.uint8
            if labels_i16 and k in self._LABELS:
                dt = np.int16
            if out is not None and k in out:
                a = out[k]
                if a.shape != (n, h, w) or a.dtype != dt or not a.flags.c_contiguous:
                    raise StereoB200Error(f"out[{k!r}]: expected a C-contiguous {np.dtype(dt).name} array of shape {(n, h, w)}")
            else:
                a = np.empty((n, h, w), dt)
            res[k] = a
            setattr(li if dt == np.int16 else o, k, a.ctypes.data)
        if labels_i16:
            self._ck(self.lib.sb200_pipeline_batch_i16(self.h, C.byref(p), _ptr(lefts), _ptr(rights), ch, w, h, n, C.byref(li),
                                                       C.byref(o)))
        else:
            self._ck(self.lib.sb200_pipeline_batch(self.h, C.byref(p), _ptr(lefts), _ptr(rights), ch, w, h, n, C.byref(o)))
        return res

    def _dev_check(self, t, name, numel, dtype_name):
        """a device tensor handed to the C ABI as a raw pointer: right device, dtype, contiguous, large enough"""
        if t is None or isinstance(t, int):
            return
        if not t.is_cuda or t.device.index != self.device:
            raise StereoB200Error(f"{name}: expected a tensor on cuda:{self.device}, got {t.device}")
        if str(t.dtype) != "torch." + dtype_name:
            raise StereoB200Error(f"{name}: expected dtype {dtype_name}, got {t.dtype}")
        if not t.is_contiguous():
            raise StereoB200Error(f"{name}: tensor must be contiguous")
        if t.numel() < numel:
            raise StereoB200Error(f"{name}: {t.numel()} elements, need {numel}")

    def _dev_outputs(self, outs, numel=0):
        o = _Outputs()
        for k, t in outs.items():
            if k not in self._F32 + self._F32_OPT + self._U8:
                raise TypeError(f"unknown output {k}")
            self._dev_check(t, k, numel, "uint8" if k in self._U8 else "float32")
            setattr(o, k, t if isinstance(t, int) else t.data_ptr())
        return o

    def pipeline_dev(self, d_left, d_right, channels, w, h, outs, params=None):
        """device tensors in, device tensors out (dict name -> tensor), asynchronous"""
        p = params or default_params()
        for name, t in (("left", d_left), ("right", d_right)):
            self._dev_check(t, name, w * h * channels, "uint8")
        o = self._dev_outputs(outs, w * h)
        self._ck(self.lib.sb200_pipeline_dev(self.h, C.byref(p), _ptr(d_left), _ptr(d_right), channels, w, h, C.byref(o)))

    def pipeline_batch_dev(self, d_left, d_right, channels, w, h, n_pairs, outs, params=None):
        p = params or default_params()
        for name, t in (("left", d_left), ("right", d_right)):
            self._dev_check(t, name, n_pairs * w * h * channels, "uint8")
        o = self._dev_outputs(outs, n_pairs * w * h)
        self._ck(self.lib.sb200_pipeline_batch_dev(self.h, C.byref(p), _ptr(d_left), _ptr(d_right), channels, w, h,
                                                   n_pairs, C.byref(o)))

    def pipeline_strip_dev(self, d_left, d_right, channels, w, strip, outs, params=None):
        """strip: dict(y0, rows, halo_top, halo_bot, frame_h)"""
        p = params or default_params()
        s = _Strip(**strip)
        held = s.halo_top + s.rows + s.halo_bot
        for name, t in (("left", d_left), ("right", d_right)):
            self._dev_check(t, name, held * w * channels, "uint8")
        o = self._dev_outputs(outs, s.rows * w)
        self._ck(self.lib.sb200_pipeline_strip_dev(self.h, C.byref(p), _ptr(d_left), _ptr(d_right), channels, w,
                                                   C.byref(s), C.byref(o)))

    def pipeline_strips_nccl(self, comm, d_own_left, d_own_right, channels, w, frame_h, y0, rows, outs, params=None):
        """this rank's rows [y0, y0+rows) of a frame_h-row frame (device tensors) -> outputs for those rows; the halo rows
        come from the neighbouring ranks over NCCL inside the call.  comm: sharding.NcclComm (or a raw ncclComm_t int)."""
        p = params or default_params()
        for name, t in (("left", d_own_left), ("right", d_own_right)):
            self._dev_check(t, name, rows * w * channels, "uint8")
        o = self._dev_outputs(outs, rows * w)
        handle, rank, world = (comm.handle, comm.rank, comm.world) if hasattr(comm, "handle") else comm
        self._ck(self.lib.sb200_pipeline_strips_nccl(self.h, C.byref(p), handle, rank, world, _ptr(d_own_left),
                                                     _ptr(d_own_right), channels, w, frame_h, y0, rows, C.byref(o)))

    def strip_rows(self, frame_h, rank, world):
        y0, rows = C.c_int(), C.c_int()
        self._ck(self.lib.sb200_strip_rows(frame_h, rank, world, C.byref(y0), C.byref(rows)))
        return y0.value, rows.value

    def strip_halo_rows(self, params=None):
        p = params or default_params()
        return self.lib.sb200_strip_halo_rows(C.byref(p))

    def view_volume_dev(self, d_guide, d_other, w, h, dmin, size_d, d_volume, d_best=None, d_disp=None, params=None):
        """the fused view kernel keeping the filtered cost volume q: d_volume is (size_d, h, w) float32 on the device"""
        p = params or default_params()
        for name, t in (("guide", d_guide), ("other", d_other)):
            self._dev_check(t, name, w * h, "uint8")
        self._dev_check(d_volume, "volume", size_d * w * h, "float32")
        for name, t in (("best", d_best), ("disp", d_disp)):
            self._dev_check(t, name, w * h, "float32")
        self._ck(self.lib.sb200_view_volume_dev(self.h, C.byref(p), _ptr(d_guide), _ptr(d_other), w, h, dmin, size_d,
                                                _ptr(d_volume), _ptr(d_best), _ptr(d_disp)))

    def subpixel_refine_dev(self, d_volume, d_disp, d_out, w, h, dmin, size_d, d_occlusion=None, d_filled=None):
        """parabola through the filtered costs at label-1, label, label+1 (beyond the reference; stereo_b200.h)"""
        self._dev_check(d_volume, "volume", size_d * w * h, "float32")
        for name, t in (("disp", d_disp), ("out", d_out), ("occlusion", d_occlusion), ("filled", d_filled)):
            self._dev_check(t, name, w * h, "float32")
        self._ck(self.lib.sb200_subpixel_refine_dev(self.h, _ptr(d_volume), _ptr(d_disp), _ptr(d_occlusion), _ptr(d_filled),
                                                    _ptr(d_out), w, h, dmin, size_d))

    def view_disparity_dev(self, d_guide, d_other, w, h, dmin, size_d, d_best, d_disp, d_mean=None, params=None):
        p = params or default_params()
        for name, t, dt in (("guide", d_guide, "uint8"), ("other", d_other, "uint8"), ("best", d_best, "float32"),
                            ("disp", d_disp, "float32"), ("mean", d_mean, "uint8")):
            self._dev_check(t, name, w * h, dt)
        self._ck(self.lib.sb200_view_disparity_dev(self.h, C.byref(p), _ptr(d_guide), _ptr(d_other), w, h, dmin, size_d,
                                                   _ptr(d_best), _ptr(d_disp), _ptr(d_mean)))

    def lr_check_fill_dev(self, d_dl, d_dr, w, h, d_occlusion, v_min, d_occ, d_filled, params=None):
        p = params or default_params()
        for name, t in (("dL", d_dl), ("dR", d_dr), ("occlusion", d_occ), ("filled", d_filled)):
            self._dev_check(t, name, w * h, "float32")
        self._ck(self.lib.sb200_lr_check_fill_dev(self.h, C.byref(p), _ptr(d_dl), _ptr(d_dr), w, h, d_occlusion,
                                                  float(v_min), _ptr(d_occ), _ptr(d_filled)))

    def compute_guided_filter_dev(self, d_i, d_cost, d_filter_cost, d_disp_map, d_mean, w, h, size_d, dmin, params=None):
        p = params or default_params()
        self._dev_check(d_i, "i", w * h, "uint8")
        self._dev_check(d_cost, "cost", w * h * size_d, "float32")
        self._dev_check(d_filter_cost, "filter_cost", w * h, "float32")
        self._dev_check(d_disp_map, "disp_map", w * h, "float32")
        self._dev_check(d_mean, "mean", w * h, "uint8")
        self._ck(self.lib.sb200_compute_guided_filter_dev(self.h, C.byref(p), _ptr(d_i), _ptr(d_cost),
                                                          _ptr(d_filter_cost), _ptr(d_disp_map), _ptr(d_mean), w, h,
                                                          size_d, dmin))

    def compute_cost_dev(self, d_i1, d_i2, d_cost, w, h, dmin, params=None):
        p = params or default_params()
        self._ck(self.lib.sb200_compute_cost_dev(self.h, C.byref(p), _ptr(d_i1), _ptr(d_i2), _ptr(d_cost), w, h, dmin))
