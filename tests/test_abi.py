"""The C-ABI library loads and exports every symbol include/ndt_b200.h declares;
the struct mirror in include/ndt_abi.h matches the reference's own headers."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

import ndt_b200


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ndt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ndt_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = ndt_b200.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"libndt_b200.so does not export {n}"


def test_version_and_error_strings():
    L = ndt_b200.lib()
    assert b"sm_100a" in L.ndt_b200_version()
    assert L.ndt_b200_flat_validate(b"\0" * 16, 16) < 0
    assert b"truncated" in L.ndt_b200_last_error()


def test_struct_sizes_of_python_mirror():
    assert C.sizeof(ndt_b200.Stats) == 80
    assert C.sizeof(ndt_b200.FlatHeader) == 256


def test_header_compiles_as_c_and_cxx(tmp_path):
    for comp, ext, std in (("gcc", "c", "-std=gnu99"), ("g++", "cpp", "-std=c++17")):
        f = tmp_path / f"t.{ext}"
        f.write_text('#include "ndt_abi.h"\n#include "ndt_b200.h"\nint main(void){return (sizeof(ndt_flat_header)==256 && sizeof(ndt_b200_stats)==80 && sizeof(ndt_flat_object)==96 && sizeof(ndt_flat_node)==32 && sizeof(ndt_flat_light)==56)?0:1;}\n')
        exe = tmp_path / f"t_{ext}"
        subprocess.run([comp, std, "-I" + os.path.join(ROOT, "include"), str(f), "-o", str(exe)], check=True)
        assert subprocess.run([str(exe)]).returncode == 0


REF = "/root/reference"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "object.h")), reason="reference headers not present")
def test_abi_mirror_matches_reference_headers(tmp_path):
    """sizeof/offsetof of our mirror == the reference's own structs."""
    pairs = [
        ("vectNd", "ndtabi_vec", ["v", "n"], ["v", "n"]),
        ("bounding_sphere", "ndtabi_bsphere", ["center", "radius", "radius_sqr"], ["center", "radius", "radius_sqr"]),
        ("object", "ndtabi_object",
         ["dimensions", "red", "red_r", "refract_index", "pos", "n_pos", "dir", "n_dir", "size", "n_size",
          "flag", "n_flag", "obj", "n_obj", "bounds", "prepped", "type_name", "intersect", "get_color", "get_reflect"],
         ["dimensions", "rgb", "refl", "refract_index", "pos", "n_pos", "dir", "n_dir", "size", "n_size",
          "flag", "n_flag", "obj", "n_obj", "bounds", "prepped", "type_name", "intersect", "get_color", "get_reflect"]),
        ("light", "ndtabi_light", ["pos", "dir", "radius", "type", "red", "angle", "u1"],
         ["pos", "dir", "radius", "type", "rgb", "angle", "u1"]),
        ("camera", "ndtabi_camera", ["type", "focal_distance", "pos", "dirX", "dirY", "imgOrig", "localZ"],
         ["type", "focal_distance", "pos", "dirX", "dirY", "imgOrig", "localZ"]),
        ("scene", "ndtabi_scene", ["dimensions", "cam", "num_objects", "num_lights", "object_ptrs", "lights", "ambient", "bg_red", "name"],
         ["dimensions", "cam", "num_objects", "num_lights", "object_ptrs", "lights", "ambient", "bg", "name"]),
        ("kd_node_t", "ndtabi_kd_node", ["dim", "boundary", "num", "obj_ids", "objs", "left", "right"],
         ["dim", "boundary", "num", "obj_ids", "objs", "left", "right"]),
        ("kd_tree_t", "ndtabi_kd_tree", ["bb", "inf_obj_ptrs", "obj_num", "inf_obj_num", "root"],
         ["bb_lower", "inf_obj_ptrs", "obj_num", "inf_obj_num", "root"]),
    ]
    lines = ['#include <stddef.h>', '#include "scene.h"', '#include "kd-tree.h"', '#include "ndt_abi.h"']
    for ref_t, our_t, rf, of in pairs:
        lines.append(f"_Static_assert(sizeof({ref_t}) == sizeof({our_t}), \"size {ref_t}\");")
        for a, b in zip(rf, of):
            lines.append(f"_Static_assert(offsetof({ref_t}, {a}) == offsetof({our_t}, {b}), \"{ref_t}.{a}\");")
    lines.append("int main(void){return 0;}")
    f = tmp_path / "abi.c"
    f.write_text("\n".join(lines) + "\n")
    subprocess.run(["gcc", "-std=gnu99", "-D_GNU_SOURCE", "-I" + REF, "-I" + os.path.join(ROOT, "include"),
                    "-c", str(f), "-o", str(tmp_path / "abi.o")], check=True)
