"""ndt_b200 -- B200-native render path for the N-dimensional tracer `ndt`.

Host-side mirror of the C ABI in include/ndt_b200.h (ctypes; no compute
happens in Python).  The product path is libndt_b200.so: hand-written CUDA
for sm_100a behind plain-C entry points.  There is no CPU fallback -- every
device call raises NdtB200Error when the library or a CUDA device is missing.

    flat = ndt_b200.flatten(scene_ptr, kdtree_ptr, w, h, host_get_bounds=ptr)
    with ndt_b200.Context(device=0) as ctx:
        ctx.upload(flat)
        frame = ctx.render_tile(0, 0, w, h)      # Frame: rgba_f64, rgba_u8, hit, obj_id, inv_depth, stats
"""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NDT_B200_LIB") or os.path.join(_HERE, "libndt_b200.so")

OPT_COUNT_FLOPS = 1
OPT_FUSED = 2


class NdtB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ndt_b200 error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("rays_primary", C.c_uint64), ("rays_bounce", C.c_uint64),
                ("rays_shadow", C.c_uint64), ("rays_unique", C.c_uint64),
                ("rays_ref", C.c_uint64), ("samples", C.c_uint64),
                ("flops", C.c_uint64), ("launches", C.c_uint64),
                ("generations", C.c_uint32), ("reserved", C.c_uint32),
                ("device_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class _HostApi(C.Structure):
    _fields_ = [("object_get_bounds", C.c_void_p), ("vectNd_rotate2", C.c_void_p)]


MONO, SIDE_SIDE_3D, OVER_UNDER_3D, ANAGLYPH_3D, HIDEF_3D = range(5)      # ndt.c:46-48
CAMERA_NORMAL, CAMERA_VR, CAMERA_PANO = range(3)                          # camera.h:16-20


class FlatHeader(C.Structure):
    """include/ndt_flat.h: ndt_flat_header"""
    _fields_ = [("magic", C.c_uint32), ("version", C.c_uint32), ("total_bytes", C.c_uint64),
                ("n", C.c_int32), ("npad", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("max_optic_depth", C.c_int32), ("specular", C.c_int32),
                ("n_items", C.c_int32), ("n_objects", C.c_int32),
                ("n_nodes", C.c_int32), ("n_leaf_refs", C.c_int32), ("n_inf", C.c_int32),
                ("n_lights", C.c_int32), ("max_leaf", C.c_int32), ("tree_depth", C.c_int32),
                ("use_focal", C.c_int32), ("reserved", C.c_int32),
                ("bg", C.c_double * 4), ("ambient", C.c_double * 3), ("focal_scale", C.c_double),
                ("off_camera", C.c_uint64), ("off_aabb", C.c_uint64), ("off_objects", C.c_uint64),
                ("off_bspheres", C.c_uint64), ("off_geom", C.c_uint64), ("n_geom", C.c_uint64),
                ("off_nodes", C.c_uint64), ("off_leaf_refs", C.c_uint64), ("off_inf", C.c_uint64),
                ("off_lights", C.c_uint64),
                ("cam_type", C.c_int32), ("stereo_mode", C.c_int32), ("view_eyes", C.c_int32),
                ("aa_pad", C.c_int32), ("off_view", C.c_uint64), ("cam_dist", C.c_double)]


_lib = None


def lib():
    """Load libndt_b200.so (built in-tree by ndt_b200/csrc/Makefile).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NdtB200Error(-4, f"{LIB_PATH} not built: run `make -C ndt_b200/csrc` "
                               "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.ndt_b200_last_error.restype = C.c_char_p
    L.ndt_b200_version.restype = C.c_char_p
    L.ndt_b200_flatten.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(_HostApi), C.POINTER(C.c_void_p)]
    L.ndt_b200_flatten_view.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(_HostApi), C.POINTER(C.c_void_p)]
    L.ndt_b200_flatten_aa.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(_HostApi), C.POINTER(C.c_void_p)]
    L.ndt_b200_render_aa.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.POINTER(C.c_uint64), C.POINTER(Stats)]
    L.ndt_b200_free_flat.argtypes = [C.c_void_p]
    L.ndt_b200_free_flat.restype = None
    L.ndt_b200_flat_validate.argtypes = [C.c_void_p, C.c_size_t]
    L.ndt_b200_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.ndt_b200_destroy.argtypes = [C.c_void_p]
    L.ndt_b200_destroy.restype = None
    L.ndt_b200_upload.argtypes = [C.c_void_p, C.c_void_p]
    L.ndt_b200_set_options.argtypes = [C.c_void_p, C.c_uint32]
    if hasattr(L, "ndt_b200_set_pool"):      # absent from round-1 builds (A/B runs with NDT_B200_LIB)
        L.ndt_b200_set_pool.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
    L.ndt_b200_render_tile.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 5 + [C.POINTER(Stats)]
    L.ndt_b200_launch_tile.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 5
    L.ndt_b200_sync.argtypes = [C.c_void_p]
    L.ndt_b200_last_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.ndt_b200_stream.argtypes = [C.c_void_p]
    L.ndt_b200_stream.restype = C.c_void_p
    L.ndt_b200_kd_tree_build.argtypes = [C.c_void_p, C.c_void_p]
    L.ndt_b200_host_alloc.argtypes = [C.c_size_t]
    L.ndt_b200_host_alloc.restype = C.c_void_p
    L.ndt_b200_host_free.argtypes = [C.c_void_p]
    L.ndt_b200_host_free.restype = None
    L.ndt_b200_trace_rays.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 8
    L.ndt_b200_fp64_peak.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.ndt_b200_replay_samples.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ndt_b200_render_image.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_HostApi), C.c_char_p, C.c_char_p,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    if hasattr(L, "ndt_b200_mgpu_init"):
        L.ndt_b200_mgpu_init.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.ndt_b200_mgpu_destroy.argtypes = [C.c_void_p]
        L.ndt_b200_mgpu_destroy.restype = None
        L.ndt_b200_mgpu_devices.argtypes = [C.c_void_p]
        L.ndt_b200_mgpu_render_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.POINTER(Stats)]
        L.ndt_b200_mgpu_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ndt_b200_mgpu_wait.argtypes = [C.c_void_p, C.POINTER(Stats)]
    _lib = L
    return L


def _check(rc):
    if rc < 0:
        raise NdtB200Error(rc, lib().ndt_b200_last_error().decode(errors="replace"))
    return rc


class FlatScene:
    """A flat scene blob (include/ndt_flat.h) held as bytes in host memory."""

    def __init__(self, blob):
        self.blob = bytes(blob)
        self.header = FlatHeader.from_buffer_copy(self.blob[:C.sizeof(FlatHeader)])
        _check(lib().ndt_b200_flat_validate(self.blob, len(self.blob)))

    @classmethod
    def load(cls, path):
        import gzip
        op = gzip.open if str(path).endswith(".gz") else open
        with op(path, "rb") as f:
            return cls(f.read())

    def save(self, path):
        import gzip
        if str(path).endswith(".gz"):
            with open(path, "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0, filename="") as f:
                f.write(self.blob)        # mtime=0: byte-identical files for identical scenes
        else:
            with open(path, "wb") as f:
                f.write(self.blob)

    def retarget(self, width, height):
        """The same scene for another frame size is NOT a header edit: the camera
        basis was scaled for width/height (ndt.c:926).  Only equal aspect ratios
        can be retargeted without re-flattening."""
        h = self.header
        if h.off_view:
            raise ValueError("view tables (VR / PANO / stereo) are per column and row; flatten the host scene again")
        if width * h.height != height * h.width:
            raise ValueError("aspect ratio differs; flatten the host scene again")
        hdr = FlatHeader.from_buffer_copy(self.blob[:C.sizeof(FlatHeader)])
        hdr.width, hdr.height = width, height
        return FlatScene(bytes(hdr) + self.blob[C.sizeof(FlatHeader):])

    def __len__(self):
        return len(self.blob)


def flatten(scene_ptr, kdtree_ptr, width, height, max_optic_depth=128, specular=1, host_get_bounds=None,
            stereo_mode=MONO, host_rotate2=None):
    """ndt_b200_flatten_view: host `scene*` + `kd_tree_t*` (ndt.c:68) -> FlatScene."""
    L = lib()
    host = _HostApi(host_get_bounds, host_rotate2)
    out = C.c_void_p()
    _check(L.ndt_b200_flatten_view(scene_ptr, kdtree_ptr, width, height, max_optic_depth, specular, stereo_mode,
                                   C.byref(host), C.byref(out)))
    try:
        total = C.c_uint64.from_address(out.value + 8).value
        return FlatScene(C.string_at(out.value, total))
    finally:
        L.ndt_b200_free_flat(out)


def flatten_aa(scene_ptr, kdtree_ptr, width, height, max_optic_depth=128, specular=1, host_get_bounds=None):
    """ndt_b200_flatten_aa: the (width+1) x (height+1) corner-sample grid of recursive anti-aliasing."""
    L = lib()
    host = _HostApi(host_get_bounds, None)
    out = C.c_void_p()
    _check(L.ndt_b200_flatten_aa(scene_ptr, kdtree_ptr, width, height, max_optic_depth, specular,
                                 C.byref(host), C.byref(out)))
    try:
        total = C.c_uint64.from_address(out.value + 8).value
        return FlatScene(C.string_at(out.value, total))
    finally:
        L.ndt_b200_free_flat(out)


def kd_tree_build(kd_tree_ptr, kd_item_list_ptr):
    """ndt_b200_kd_tree_build: drop-in for kd_tree_build (kd-tree.c:421) on host structures."""
    return _check(lib().ndt_b200_kd_tree_build(kd_tree_ptr, kd_item_list_ptr))


def kd_tree_build_bounded(kd_tree_ptr, kd_item_list_ptr, max_depth=16, leaf_size=64, max_growth=1.5):
    """ndt_b200_kd_tree_build_bounded: bounded-depth variant for scenes the reference's builder cannot finish."""
    L = lib()
    L.ndt_b200_kd_tree_build_bounded.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double]
    return _check(L.ndt_b200_kd_tree_build_bounded(kd_tree_ptr, kd_item_list_ptr, max_depth, leaf_size, max_growth))


def pinned_empty(shape, dtype):
    """numpy array over page-locked host memory (ndt_b200_host_alloc); freed with the array."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    p = lib().ndt_b200_host_alloc(n)
    if not p:
        raise NdtB200Error(-3, lib().ndt_b200_last_error().decode(errors="replace"))
    buf = (C.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    import weakref
    weakref.finalize(buf, lib().ndt_b200_host_free, p)
    return arr


class Frame:
    def __init__(self, tw, th, want, pinned=False):
        mk = pinned_empty if pinned else (lambda shape, dt: np.empty(shape, dt))
        self.rgba_f64 = mk((th, tw, 4), np.float64) if "f64" in want else None
        self.rgba_u8 = mk((th, tw, 4), np.uint8) if "u8" in want else None
        self.hit = mk((th, tw), np.uint8) if "hit" in want else None
        self.obj_id = mk((th, tw), np.int32) if "id" in want else None
        self.inv_depth = mk((th, tw), np.float64) if "depth" in want else None
        self.stats = None


def _p(a):
    return a.ctypes.data if a is not None else None


class Context:
    """One per GPU (ndt_b200_ctx).  Not re-entrant."""

    ALL = ("f64", "u8", "hit", "id", "depth")

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().ndt_b200_init(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            lib().ndt_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            if sys is None or sys.is_finalizing():
                return  # the CUDA runtime may already be gone at interpreter shutdown
            self.close()
        except Exception:
            pass

    def set_options(self, options):
        _check(lib().ndt_b200_set_options(self._h, options))

    def set_pool(self, bounce_factor=6.0, slack_records=65536, rays_per_batch=0):
        """ndt_b200_set_pool: record pool = n0 * (1 + bounce_factor) + slack_records, batches of rays_per_batch."""
        _check(lib().ndt_b200_set_pool(self._h, bounce_factor, slack_records, rays_per_batch))

    def upload(self, flat):
        self._flat = flat  # keep the bytes alive until the async copy is consumed
        _check(lib().ndt_b200_upload(self._h, flat.blob))

    def render_tile(self, x0, y0, tw, th, want=ALL, out=None):
        """ndt_b200_render_tile: host buffers, device->host copies included."""
        fr = out or Frame(tw, th, want)
        st = Stats()
        _check(lib().ndt_b200_render_tile(self._h, x0, y0, tw, th, _p(fr.rgba_f64), _p(fr.rgba_u8),
                                          _p(fr.hit), _p(fr.obj_id), _p(fr.inv_depth), C.byref(st)))
        fr.stats = st
        return fr

    def render_aa(self, aa_diff=20, aa_depth=4, want_f64=False):
        """ndt_b200_render_aa on a scene from flatten_aa: (u8 [H, W, 4], f64 or None, pixels resampled, stats)."""
        h = self._flat.header
        W, H = h.width - 1, h.height - 1
        u8 = np.empty((H, W, 4), np.uint8)
        f64 = np.empty((H, W, 4), np.float64) if want_f64 else None
        n = C.c_uint64(0)
        st = Stats()
        _check(lib().ndt_b200_render_aa(self._h, aa_diff, aa_depth, u8.ctypes.data, _p(f64), C.byref(n), C.byref(st)))
        return u8, f64, n.value, st

    def launch_tile(self, x0, y0, tw, th, d_f64=None, d_u8=None, d_hit=None, d_id=None, d_depth=None):
        """ndt_b200_launch_tile: outputs are DEVICE pointers (ints, e.g. tensor.data_ptr())."""
        _check(lib().ndt_b200_launch_tile(self._h, x0, y0, tw, th, d_f64, d_u8, d_hit, d_id, d_depth))

    def sync(self):
        _check(lib().ndt_b200_sync(self._h))
        st = Stats()
        _check(lib().ndt_b200_last_stats(self._h, C.byref(st)))
        return st

    @property
    def stream(self):
        return lib().ndt_b200_stream(self._h)

    def trace_rays(self, origins, dirs, dist_limits=None):
        """ndt_b200_trace_rays: returns (found, obj_id, t, hit, normal) numpy arrays."""
        o = np.ascontiguousarray(origins, np.float64)
        v = np.ascontiguousarray(dirs, np.float64)
        n = o.shape[0]
        lim = None if dist_limits is None else np.ascontiguousarray(dist_limits, np.float64)
        found = np.zeros(n, np.int32); oid = np.zeros(n, np.int32); t = np.zeros(n, np.float64)
        hit = np.zeros_like(o); nrm = np.zeros_like(o)
        _check(lib().ndt_b200_trace_rays(self._h, n, o.ctypes.data, v.ctypes.data, _p(lim), found.ctypes.data,
                                         oid.ctypes.data, t.ctypes.data, hit.ctypes.data, nrm.ctypes.data))
        return found, oid, t, hit, nrm

    def replay_samples(self, rgba):
        """ndt_b200_replay_samples: (averaged colours [n,4], samples taken [n]) for traced colours [n,4]"""
        a = np.ascontiguousarray(rgba, np.float64).reshape(-1, 4)
        out = np.zeros_like(a); ns = np.zeros(len(a), np.int32)
        _check(lib().ndt_b200_replay_samples(self._h, len(a), a.ctypes.data, out.ctypes.data, ns.ctypes.data))
        return out, ns

    def fp64_peak(self, fused):
        g = C.c_double(0)
        _check(lib().ndt_b200_fp64_peak(self._h, 1 if fused else 0, C.byref(g)))
        return g.value


class MultiGpu:
    """ndt_b200_mgpu_*: all (or the first n) GPUs of the box in ONE process -- a frame as row bands, or an
    animation frame by frame -- gathered in host buffers."""

    def __init__(self, n_devices=0):
        self._h = C.c_void_p()
        _check(lib().ndt_b200_mgpu_init(n_devices, None, C.byref(self._h)))
        self._keep = []

    @property
    def devices(self):
        return lib().ndt_b200_mgpu_devices(self._h)

    def close(self):
        if self._h:
            lib().ndt_b200_mgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def render_frame(self, flat, band_rows=0, want=Context.ALL, out=None):
        h = flat.header
        fr = out or Frame(h.width, h.height, want)
        st = Stats()
        _check(lib().ndt_b200_mgpu_render_frame(self._h, flat.blob, band_rows, _p(fr.rgba_f64), _p(fr.rgba_u8),
                                                _p(fr.hit), _p(fr.obj_id), _p(fr.inv_depth), C.byref(st)))
        fr.stats = st
        return fr

    def submit(self, flat, u8=None, f64=None):
        """queue one frame of an animation; u8 / f64 are numpy arrays that receive it (kept alive until wait())"""
        self._keep.append((u8, f64))
        _check(lib().ndt_b200_mgpu_submit(self._h, flat.blob, _p(u8), _p(f64)))

    def wait(self):
        st = Stats()
        try:
            _check(lib().ndt_b200_mgpu_wait(self._h, C.byref(st)))
        finally:
            self._keep.clear()
        return st
