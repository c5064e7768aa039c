import sys, time, ctypes as C
sys.path.insert(0,'/root/repo')
from oracle.refharness import RefHarness, rgba_f64_to_u8
import numpy as np
R = RefHarness()
F = C.CDLL('/tmp/libflat_test.so')
O = C.CDLL('/root/repo/oracle/libndt_oracle.so')
F.ndt_b200_last_error.restype = C.c_char_p
class Host(C.Structure): _fields_=[('get_bounds', C.c_void_p)]
F.ndt_b200_flatten.argtypes=[C.c_void_p,C.c_void_p,C.c_int,C.c_int,C.c_int,C.c_int,C.POINTER(Host),C.POINTER(C.c_void_p)]
O.ndo_render.argtypes=[C.c_void_p]+[C.c_int]*5+[C.c_void_p]*6
def run(dims, scene, w, h, cfg=None, frame=0, mod=128):
    R.open_scene(scene)
    fr = R.scene_frames(dims, cfg) if scene else 300
    R.begin_frame(dims, frame, fr, cfg)
    hst = Host(R.get_bounds_ptr); out = C.c_void_p()
    r = F.ndt_b200_flatten(R.scene_ptr, R.kdtree_ptr, w,h,mod,1, C.byref(hst), C.byref(out))
    assert r==0, F.ndt_b200_last_error()
    img, s = R.render(w,h,threads=8,max_optic_depth=mod)
    hit, oid, dist = R.primary(w,h)
    R.end_frame()
    rgba = np.zeros((h,w,4)); u8=np.zeros((h,w,4),np.uint8); ohit=np.zeros((h,w),np.uint8); oid2=np.zeros((h,w),np.int32); dep=np.zeros((h,w))
    st=(C.c_uint64*5)()
    t=time.time()
    O.ndo_render(out, 0,0,w,h,8, rgba.ctypes.data,u8.ctypes.data,ohit.ctypes.data,oid2.ctypes.data,dep.ctypes.data, st)
    t=time.time()-t
    same = (rgba.view(np.uint64)==img.view(np.uint64))
    print(f"{scene} d={dims} cfg={cfg} f={frame} {w}x{h}: ref {s:.2f}s port {t:.2f}s | f64 bit-identical px {same.all(axis=2).mean()*100:.4f}% maxabs {np.nanmax(np.abs(rgba-img)):.3g} | hit mism {(ohit!=hit).sum()} id mism {(oid2!=oid).sum()} | rays {list(st)}")
    return rgba, img
if __name__=="__main__":
    run(4,None,192,108)
    run(3,None,96,54)
    run(5,None,96,54, frame=37)
    run(5,'balls',96,54, frame=2)
    run(10,'mixed10d',96,54)
    run(7,'mixed10d',96,54, frame=5)
    run(6,'hypercube-points',96,54)
    run(5,'hypercube',64,36,'hcube')
    run(8,'hypercube',96,54)
