"""ctypes driver for oracle/_ref (the unmodified reference) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  It loads oracle/_ref/libndt_ref.so
(built by oracle/Makefile from the sources under /root/reference) with
RTLD_GLOBAL so that the reference's object/scene plugins resolve the host
symbols they import, and walks one frame through the
same prologue main() uses (ndt.c:1791-1933).
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REFDIR, "libndt_ref.so"))


class RefHarness:
    """One process-wide instance per library flavour (plain / ray-counting)."""

    _inst = {}

    def __new__(cls, counting=False):
        key = bool(counting)
        if key in cls._inst:
            return cls._inst[key]
        if cls._inst:
            # the two flavours export the same global symbols; mixing them in
            # one process would cross-wire the plugins.
            raise RuntimeError("only one flavour of the reference library per process")
        self = super().__new__(cls)
        name = "libndt_ref_count.so" if counting else "libndt_ref.so"
        path = os.path.join(REFDIR, name)
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L = self.lib
        L.refh_init.argtypes = [C.c_char_p]
        L.refh_open_scene.argtypes = [C.c_char_p]
        L.refh_scene_frames.argtypes = [C.c_int, C.c_char_p]
        L.refh_skip_frame.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.refh_begin_frame.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_double)]
        L.refh_scene.restype = C.c_void_p
        L.refh_kdtree.restype = C.c_void_p
        L.refh_object_get_bounds_ptr.restype = C.c_void_p
        L.refh_render.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
        L.refh_render_ex.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
        L.refh_render_aa.argtypes = [C.c_int] * 6 + [C.c_void_p, C.POINTER(C.c_double)]
        L.refh_write_yaml.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
        L.refh_begin_frame_nokd.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.refh_trace_brute.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.POINTER(C.c_int)]
        L.refh_set_camera.argtypes = [C.c_int, C.c_double, C.c_double]
        L.refh_rotate2_ptr.restype = C.c_void_p
        L.refh_primary.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refh_trace_ray.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.refh_ray_count.restype = C.c_long
        L.refh_ray_count.argtypes = [C.c_int]
        L.refh_set_specular.argtypes = [C.c_int]
        L.refh_items.restype = C.c_void_p
        L.refh_alt_tree_begin.restype = C.c_void_p
        L.refh_alt_tree_build_reference.restype = C.c_double
        L.refh_init(os.path.join(REFDIR, "objects").encode())
        self.counting = key
        self._scene = None
        self._next_frame = 0
        self.kd_seconds = 0.0
        cls._inst[key] = self
        return self

    # -- scene / frame management -------------------------------------------
    def open_scene(self, scene):
        """scene: None/'' for the built-in test scene, else a plugin name like 'hypercube'."""
        path = b"" if not scene else os.path.join(REFDIR, "scenes", scene + ".so").encode()
        if self.lib.refh_open_scene(path) != 0:
            raise RuntimeError(f"cannot open scene {scene!r}")
        self._scene = scene or ""
        self._next_frame = 0

    def _cfg(self, cfg):
        """scenes/yaml.c takes a file name as its config string: repo-relative names are resolved here"""
        if cfg and self._scene == "yaml" and not os.path.isabs(cfg):
            cfg = os.path.join(os.path.dirname(HERE), cfg)
        return cfg

    def scene_frames(self, dims, cfg=None):
        cfg = self._cfg(cfg)
        return self.lib.refh_scene_frames(dims, cfg.encode() if cfg else None)

    def begin_frame(self, dims, frame, frames, cfg=None):
        """scene_setup(frame) + kd build + camera_aim.  Stateful scenes (balls.c)
        need every earlier frame's scene_setup to have run, in order."""
        cfg = self._cfg(cfg)
        cfgb = cfg.encode() if cfg else None
        if frame < self._next_frame:
            self.open_scene(self._scene)  # restart the plugin's state
        for f in range(self._next_frame, frame):
            self.lib.refh_skip_frame(dims, f, frames, cfgb)
        kd = C.c_double(0)
        r = self.lib.refh_begin_frame(dims, frame, frames, cfgb, C.byref(kd))
        if r != 0:
            raise RuntimeError(f"refh_begin_frame -> {r}")
        self._next_frame = frame + 1
        self.kd_seconds = kd.value
        self.dims = dims

    def write_yaml(self, dims, frame, frames, fname, cfg=None, to_buffer=False):
        """scene_setup(frame) + the reference's scene_write_yaml / scene_write_yaml_buffer (ndt -y, ndt.c:1798-1808)"""
        if frame < self._next_frame:
            self.open_scene(self._scene)
        for f in range(self._next_frame, frame):
            self.lib.refh_skip_frame(dims, f, frames, cfg.encode() if cfg else None)
        r = self.lib.refh_write_yaml(dims, frame, frames, cfg.encode() if cfg else None, fname.encode(), int(to_buffer))
        self._next_frame = frame + 1
        if r != 0:
            raise RuntimeError(f"refh_write_yaml -> {r}")

    def begin_frame_nokd(self, dims, frame, frames, cfg=None):
        """begin_frame without kd_tree_build: the global tree is left for ndt_b200.kd_tree_build_bounded."""
        r = self.lib.refh_begin_frame_nokd(dims, frame, frames, cfg.encode() if cfg else None)
        if r != 0:
            raise RuntimeError(f"refh_begin_frame_nokd -> {r}")

    def trace_brute(self, o, v, dist_limit=-1.0):
        """trace() over every kd item without a tree (object.c:692): (return value, hit point, item id)."""
        o = np.ascontiguousarray(o, np.float64); v = np.ascontiguousarray(v, np.float64)
        hit = np.zeros_like(o); oid = C.c_int(-1)
        r = self.lib.refh_trace_brute(o.ctypes.data, v.ctypes.data, dist_limit, hit.ctypes.data, C.byref(oid))
        return r, hit, oid.value

    def end_frame(self):
        self.lib.refh_end_frame()

    def reaim(self):
        """restore the aimed camera basis, so render() can be repeated on the open frame"""
        if self.lib.refh_reaim() != 0:
            raise RuntimeError("refh_reaim: no open frame")

    @property
    def scene_ptr(self):
        return self.lib.refh_scene()

    @property
    def kdtree_ptr(self):
        return self.lib.refh_kdtree()

    @property
    def get_bounds_ptr(self):
        return self.lib.refh_object_get_bounds_ptr()

    @property
    def items_ptr(self):
        """&kditems of the open frame (kd_item_list_t*, ndt.c:1900)"""
        return self.lib.refh_items()

    def alt_tree_begin(self):
        """a fresh kd_tree_init'ed kd_tree_t over the same items, for an external builder"""
        return self.lib.refh_alt_tree_begin()

    def alt_tree_end(self):
        self.lib.refh_alt_tree_end()

    def num_items(self):
        return self.lib.refh_num_items()

    # -- the reference's results ---------------------------------------------
    def render(self, w, h, threads=None, max_optic_depth=128, stereo=0):
        """fp64 RGBA [h, w, 4] from the reference's render_image (any stereo_mode of ndt.c:46-48), and seconds."""
        threads = threads or os.cpu_count()
        out = np.empty((h, w, 4), dtype=np.float64)
        sec = C.c_double(0)
        r = self.lib.refh_render_ex(w, h, threads, max_optic_depth, stereo, out.ctypes.data, C.byref(sec))
        if r != 0:
            raise RuntimeError(f"refh_render_ex -> {r}")
        return out, sec.value

    def render_aa(self, w, h, aa_diff=20, aa_depth=4, threads=None, max_optic_depth=128):
        """8-bit RGBA [h, w, 4]: render_image with recursive_aa set (-w / -a), i.e. Whitted resampling."""
        threads = threads or os.cpu_count()
        out = np.empty((h, w, 4), dtype=np.uint8)
        sec = C.c_double(0)
        r = self.lib.refh_render_aa(w, h, threads, max_optic_depth, aa_diff, aa_depth, out.ctypes.data, C.byref(sec))
        if r != 0:
            raise RuntimeError(f"refh_render_aa -> {r}")
        return out, sec.value

    def set_camera(self, cam_type, h_fov=0.0, v_fov=0.0):
        """-V / -P of the reference's command line (ndt.c:1915-1925): camera type + fields of view, re-aimed."""
        r = self.lib.refh_set_camera(cam_type, h_fov, v_fov)
        if r != 0:
            raise RuntimeError(f"refh_set_camera -> {r}")

    @property
    def rotate2_ptr(self):
        return self.lib.refh_rotate2_ptr()

    def primary(self, w, h):
        hit = np.empty((h, w), dtype=np.uint8)
        oid = np.empty((h, w), dtype=np.int32)
        dist = np.empty((h, w), dtype=np.float64)
        r = self.lib.refh_primary(w, h, hit.ctypes.data, oid.ctypes.data, dist.ctypes.data)
        if r != 0:
            raise RuntimeError(f"refh_primary -> {r}")
        return hit, oid, dist

    def trace_ray(self, o, v, dist_limit=-1.0):
        o = np.ascontiguousarray(o, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        hit = np.zeros_like(o)
        nrm = np.zeros_like(o)
        oid = C.c_int(-1)
        r = self.lib.refh_trace_ray(o.ctypes.data, v.ctypes.data, dist_limit,
                                    hit.ctypes.data, nrm.ctypes.data, C.byref(oid))
        return r, hit, nrm, oid.value

    def ray_count(self, reset=False):
        return self.lib.refh_ray_count(1 if reset else 0)


def rgba_f64_to_u8(img):
    """image.h:36-39 pixel_d2c: (unsigned char)(sqrt(clamp(v,0,1))*255), truncating."""
    v = np.where(img < 1.0, img, 1.0)          # MIN(1.0, d): d unless 1.0 < d
    v = np.where(v < 0.0, 0.0, v)              # MAX(0.0, .)
    v = np.nan_to_num(v, nan=0.0)
    return (np.sqrt(v) * 255.0).astype(np.uint8)
