"""ctypes bindings to oracle/liboracle.so (the CPU restatement) and oracle/_ref/libref.so
(the unmodified reference sources compiled in place).  TEST INFRASTRUCTURE ONLY: imported
by tests/, __graft_entry__.smoke() and bench.py's CPU legs - never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

BOX_FAITHFUL, BOX_EXACT = 0, 1

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


class SoParams(C.Structure):
    _fields_ = [
        ("radius", C.c_int),
        ("eps", C.c_double),
        ("alpha", C.c_float),
        ("th_color", C.c_float),
        ("th_grad", C.c_float),
        ("d_lr", C.c_int),
        ("box_mode", C.c_int),
        ("use_fma", C.c_int),
        ("nthreads", C.c_int),
    ]


def build_oracle(ref=False):
    """Compile the checker (gcc) and, if asked and possible, the reference lib (nvcc)."""
    target = ["all"] if ref else ["liboracle.so"]
    subprocess.run(["make", "-s", "-C", ORACLE_DIR] + target, check=True)


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = L = C.CDLL(path)
        L.so_default_params.argtypes = [C.POINTER(SoParams)]
        L.so_max_threads.restype = C.c_int
        L.so_best_init.restype = C.c_float
        L.so_rgb_to_gray.argtypes = [u8p, u8p, C.c_int, C.c_int]
        L.so_x_derivative.argtypes = [u8p, f32p, C.c_int, C.c_int]
        L.so_cost_volume.argtypes = [C.POINTER(SoParams), u8p, u8p, f32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.so_integral.argtypes = [f32p, f32p, C.c_int, C.c_int]
        L.so_box_from_sat.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int]
        L.so_box_exact.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int]
        L.so_box_mean.argtypes = [C.POINTER(SoParams), f32p, f32p, C.c_int, C.c_int]
        L.so_guide_stats.argtypes = [C.POINTER(SoParams), u8p, f32p, f32p, f32p, C.c_void_p, C.c_int, C.c_int]
        L.so_guided_slice.argtypes = [C.POINTER(SoParams), f32p, f32p, f32p, f32p, f32p, f32p, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.so_disp_select.argtypes = [f32p, f32p, f32p, C.c_void_p, C.c_size_t, C.c_int]
        L.so_guided_filter.argtypes = [C.POINTER(SoParams), u8p, f32p, f32p, f32p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_int, C.c_int, C.c_int]
        L.so_view_disparity.argtypes = [C.POINTER(SoParams), u8p, u8p, f32p, f32p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_int, C.c_int, C.c_int]
        L.so_detect_occlusion.argtypes = [C.POINTER(SoParams), f32p, f32p, C.c_int, C.c_int, C.c_int]
        L.so_fill_occlusion.argtypes = [f32p, C.c_int, C.c_int, C.c_float]
        L.so_pipeline_gray.argtypes = [C.POINTER(SoParams), u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int] + [f32p] * 6 + [
            C.c_void_p] * 4
        L.so_write_mat.argtypes = [f32p, u8p, C.c_int, C.c_int]
        L.so_subpixel_refine.argtypes = [f32p, f32p, C.c_void_p, C.c_void_p, f32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.so_weighted_median.argtypes = [u8p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        L.so_view_disparity_rgb.argtypes = [C.POINTER(SoParams), u8p, C.c_int, u8p, u8p, f32p, f32p, C.c_void_p, C.c_int,
                                            C.c_int, C.c_int, C.c_int]

    def params(self, box_mode=BOX_FAITHFUL, use_fma=0, nthreads=1, **kw):
        p = SoParams()
        self.lib.so_default_params(C.byref(p))
        p.box_mode, p.use_fma, p.nthreads = box_mode, use_fma, nthreads
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def max_threads(self):
        return self.lib.so_max_threads()

    def best_init(self):
        return np.float32(self.lib.so_best_init())

    def rgb_to_gray(self, rgb):
        h, w, ch = rgb.shape
        out = np.empty((h, w), np.uint8)
        self.lib.so_rgb_to_gray(np.ascontiguousarray(rgb), out, h * w, ch)
        return out

    def x_derivative(self, img):
        h, w = img.shape
        out = np.empty((h, w), np.float32)
        self.lib.so_x_derivative(np.ascontiguousarray(img), out, w, h)
        return out

    def cost_volume(self, i1, i2, size_d, dmin, p=None):
        p = p or self.params()
        h, w = i1.shape
        out = np.empty((size_d, h, w), np.float32)
        self.lib.so_cost_volume(C.byref(p), np.ascontiguousarray(i1), np.ascontiguousarray(i2), out, w, h, size_d, dmin)
        return out

    def integral(self, img):
        h, w = img.shape
        out = np.empty((h, w), np.float32)
        self.lib.so_integral(np.ascontiguousarray(img, np.float32), out, w, h)
        return out

    def box_from_sat(self, sat, r=9):
        h, w = sat.shape
        out = np.empty((h, w), np.float32)
        self.lib.so_box_from_sat(np.ascontiguousarray(sat), out, w, h, r)
        return out

    def box_mean(self, img, p=None):
        p = p or self.params()
        h, w = img.shape
        out = np.empty((h, w), np.float32)
        self.lib.so_box_mean(C.byref(p), np.ascontiguousarray(img, np.float32), out, w, h)
        return out

    def guide_stats(self, img, p=None):
        p = p or self.params()
        h, w = img.shape
        I, m, v = (np.empty((h, w), np.float32) for _ in range(3))
        mu8 = np.empty((h, w), np.uint8)
        self.lib.so_guide_stats(C.byref(p), np.ascontiguousarray(img), I, m, v, _opt(mu8), w, h)
        return I, m, v, mu8

    def guided_slice(self, I, mean_I, var_I, pk, p=None):
        """returns q, a, b, mean_p, mean_Ip for one slice"""
        p = p or self.params()
        h, w = I.shape
        q, a, b, mp, mIp = (np.empty((h, w), np.float32) for _ in range(5))
        scratch = np.empty(6 * h * w, np.float32)
        self.lib.so_guided_slice(C.byref(p), I, mean_I, var_I, np.ascontiguousarray(pk), q, scratch, w, h, _opt(a),
                                 _opt(b), _opt(mp), _opt(mIp))
        return q, a, b, mp, mIp

    def disp_select(self, q, best, dmap, label, second=None):
        self.lib.so_disp_select(q, best, dmap, _opt(second), q.size, label)

    def guided_filter(self, img, cost, dmin, p=None, best=None, dmap=None, want_second=False):
        p = p or self.params()
        size_d, h, w = cost.shape
        best = np.full((h, w), self.best_init(), np.float32) if best is None else best
        dmap = np.zeros((h, w), np.float32) if dmap is None else dmap
        mu8 = np.empty((h, w), np.uint8)
        second = np.full((h, w), self.best_init(), np.float32) if want_second else None
        self.lib.so_guided_filter(C.byref(p), np.ascontiguousarray(img), np.ascontiguousarray(cost), best, dmap,
                                  _opt(mu8), _opt(second), w, h, size_d, dmin)
        return best, dmap, mu8, second

    def view_disparity(self, guide, other, size_d, dmin, p=None, want_second=False):
        p = p or self.params()
        h, w = guide.shape
        best = np.full((h, w), self.best_init(), np.float32)
        dmap = np.zeros((h, w), np.float32)
        mu8 = np.empty((h, w), np.uint8)
        second = np.full((h, w), self.best_init(), np.float32) if want_second else None
        self.lib.so_view_disparity(C.byref(p), np.ascontiguousarray(guide), np.ascontiguousarray(other), best, dmap,
                                   _opt(mu8), _opt(second), w, h, size_d, dmin)
        return best, dmap, mu8, second

    def view_disparity_rgb(self, guide_rgb, gray_guide, gray_other, size_d, dmin, p=None, want_second=False):
        """RGB-guide view (not in the reference)"""
        p = p or self.params()
        h, w, ch = guide_rgb.shape
        best = np.full((h, w), self.best_init(), np.float32)
        dmap = np.zeros((h, w), np.float32)
        second = np.full((h, w), self.best_init(), np.float32) if want_second else None
        self.lib.so_view_disparity_rgb(C.byref(p), np.ascontiguousarray(guide_rgb), ch, np.ascontiguousarray(gray_guide),
                                       np.ascontiguousarray(gray_other), best, dmap, _opt(second), w, h, size_d, dmin)
        return best, dmap, second

    def detect_occlusion(self, dL, dR, d_occlusion, p=None):
        p = p or self.params()
        h, w = dL.shape
        out = np.array(dL, np.float32, copy=True)
        self.lib.so_detect_occlusion(C.byref(p), out, np.ascontiguousarray(dR, np.float32), d_occlusion, w, h)
        return out

    def fill_occlusion(self, disp, vmin):
        h, w = disp.shape
        out = np.array(disp, np.float32, copy=True)
        self.lib.so_fill_occlusion(out, w, h, float(vmin))
        return out

    def pipeline_gray(self, gl, gr, dmin, size_d, p=None, want_second=False):
        p = p or self.params()
        h, w = gl.shape
        outs = [np.empty((h, w), np.float32) for _ in range(6)]  # dL dR occ filled bestL bestR
        mL, mR = np.empty((h, w), np.uint8), np.empty((h, w), np.uint8)
        sL = np.empty((h, w), np.float32) if want_second else None
        sR = np.empty((h, w), np.float32) if want_second else None
        self.lib.so_pipeline_gray(C.byref(p), np.ascontiguousarray(gl), np.ascontiguousarray(gr), w, h, dmin, size_d,
                                  *outs, _opt(mL), _opt(mR), _opt(sL), _opt(sR))
        keys = ["dL", "dR", "occ", "filled", "bestL", "bestR"]
        res = dict(zip(keys, outs))
        res.update(meanL=mL, meanR=mR, secondL=sL, secondR=sR)
        return res

    def weighted_median(self, gray, occlusion, filled, dmin, size_d, radius=19, sigma_space=9.0, sigma_color=25.5, nthreads=0):
        h, w = gray.shape
        g = np.ascontiguousarray(gray, np.uint8)
        oc = np.ascontiguousarray(occlusion, np.float32)
        fl = np.ascontiguousarray(filled, np.float32)
        out = np.empty((h, w), np.float32)
        self.lib.so_weighted_median(g, oc, fl, out, w, h, int(dmin), int(size_d), int(radius), float(sigma_space),
                                    float(sigma_color), int(nthreads))
        return out

    def subpixel_refine(self, vol, disp, dmin, occlusion=None, filled=None):
        size_d, h, w = vol.shape
        out = np.empty((h, w), np.float32)
        self.lib.so_subpixel_refine(np.ascontiguousarray(vol, np.float32), np.ascontiguousarray(disp, np.float32),
                                    _opt(occlusion), _opt(filled), out, w, h, dmin, size_d)
        return out

    def write_mat(self, mat):
        h, w = mat.shape
        out = np.empty((h, w), np.uint8)
        self.lib.so_write_mat(np.ascontiguousarray(mat, np.float32), out, w, h)
        return out


class RefLib:
    """oracle/_ref/libref.so: the reference's own functions (see oracle/ref_shim.cu)."""

    def __init__(self, path):
        self.lib = L = C.CDLL(path)
        L.ref_rgb_to_gray_cpu.argtypes = [u8p, u8p, C.c_int, C.c_int]
        L.ref_x_derivative_cpu.argtypes = [u8p, f32p, C.c_int, C.c_int]
        L.ref_cost_volume_cpu.argtypes = [u8p, u8p, f32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_integral_cpu.argtypes = [f32p, f32p, C.c_int, C.c_int]
        L.ref_box_filter_cpu.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int]
        L.ref_disp_select_cpu.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int]
        L.ref_guided_filter_cpu.argtypes = [u8p, f32p, f32p, f32p, u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_fill_occlusion_cpu.argtypes = [f32p, C.c_int, C.c_int, C.c_float]
        L.ref_detect_occlusion_cpu.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_view_disparity_cpu.argtypes = [u8p, u8p, f32p, f32p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_pipeline_gray_cpu.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int] + [f32p] * 6 + [C.c_int]
        L.ref_rgb_to_gray_gpu.argtypes = [u8p, u8p, C.c_int, C.c_int]
        L.ref_compute_cost_gpu.argtypes = [u8p, u8p, f32p, C.c_int, C.c_int, C.c_int]
        L.ref_integral_gpu.argtypes = [f32p, f32p, C.c_int, C.c_int]
        L.ref_compute_guided_filter_gpu.argtypes = [u8p, f32p, f32p, f32p, u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_detect_occlusion_gpu.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int]
        L.ref_fill_occlusion_gpu.argtypes = [f32p, C.c_int, C.c_int, C.c_float]
        L.ref_max_threads.restype = C.c_int
        self.size_d_macro = L.ref_size_d_macro()
        self.dmin_macro = L.ref_dmin_macro()

    def max_threads(self):
        return self.lib.ref_max_threads()

    def rgb_to_gray_cpu(self, rgb):
        h, w, ch = rgb.shape
        out = np.empty((h, w), np.uint8)
        self.lib.ref_rgb_to_gray_cpu(np.ascontiguousarray(rgb), out, h * w, ch)
        return out

    def x_derivative_cpu(self, img):
        h, w = img.shape
        out = np.empty((h, w), np.float32)
        self.lib.ref_x_derivative_cpu(np.ascontiguousarray(img), out, w, h)
        return out

    def cost_volume_cpu(self, i1, i2, size_d, dmin):
        h, w = i1.shape
        out = np.empty((size_d, h, w), np.float32)
        self.lib.ref_cost_volume_cpu(np.ascontiguousarray(i1), np.ascontiguousarray(i2), out, w, h, size_d, dmin)
        return out

    def integral_cpu(self, img):
        h, w = img.shape
        out = np.empty((h, w), np.float32)
        self.lib.ref_integral_cpu(np.ascontiguousarray(img, np.float32), out, w, h)
        return out

    def box_filter_cpu(self, img, sat):
        h, w = img.shape
        out = np.empty((h, w), np.float32)
        self.lib.ref_box_filter_cpu(np.ascontiguousarray(img, np.float32), np.ascontiguousarray(sat), out, w, h)
        return out

    def guided_filter_cpu(self, img, cost, dmin, best_init):
        size_d, h, w = cost.shape
        best = np.full((h, w), best_init, np.float32)
        dmap = np.zeros((h, w), np.float32)
        mean = np.empty((h, w), np.uint8)
        self.lib.ref_guided_filter_cpu(np.ascontiguousarray(img), np.ascontiguousarray(cost), best, dmap, mean, w, h,
                                       size_d, dmin)
        return best, dmap, mean

    def fill_occlusion_cpu(self, disp, vmin):
        h, w = disp.shape
        out = np.array(disp, np.float32, copy=True)
        self.lib.ref_fill_occlusion_cpu(out, w, h, float(vmin))
        return out

    def detect_occlusion_cpu(self, dL, dR, d_occlusion, slack=1024):
        h, w = dL.shape
        out = np.array(dL, np.float32, copy=True)
        self.lib.ref_detect_occlusion_cpu(out, np.ascontiguousarray(dR, np.float32), d_occlusion, w, h, slack)
        return out

    def view_disparity_cpu(self, guide, other, size_d, dmin, best_init, nthreads=1):
        h, w = guide.shape
        best = np.full((h, w), best_init, np.float32)
        dmap = np.zeros((h, w), np.float32)
        mean = np.empty((h, w), np.uint8)
        self.lib.ref_view_disparity_cpu(np.ascontiguousarray(guide), np.ascontiguousarray(other), best, dmap,
                                        mean.ctypes.data_as(C.c_void_p), w, h, size_d, dmin, nthreads)
        return best, dmap, mean

    def pipeline_gray_cpu(self, gl, gr, dmin, size_d, nthreads=1):
        h, w = gl.shape
        outs = [np.empty((h, w), np.float32) for _ in range(6)]
        self.lib.ref_pipeline_gray_cpu(np.ascontiguousarray(gl), np.ascontiguousarray(gr), w, h, dmin, size_d, *outs,
                                       nthreads)
        return outs  # dL dR occ filled bestL bestR


_ORACLE = None
_REF = None


def load_oracle():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE


def ref_available():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libref.so"))


def load_ref():
    global _REF
    if _REF is None:
        _REF = RefLib(os.path.join(ORACLE_DIR, "_ref", "libref.so"))
    return _REF


def load_png(name):
    from PIL import Image

    return np.array(Image.open(os.path.join(GOLDEN_DIR, "tsukuba", name)))


def tsukuba_rgb():
    return load_png("tsukuba0.png")[..., :3].copy(), load_png("tsukuba1.png")[..., :3].copy()
