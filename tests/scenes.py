"""Scene catalogue shared by the tests, the golden generator and bench.py.

Each entry names a reference scene plugin (oracle/_ref/scenes/<name>.so; None =
the built-in scene of scene.c:429-571), the dimension, the scene's config
string (-u), the frame, and a SMALL frame size at which the CPU oracle
finishes in about a second.  `cfg` rows map to BASELINE.json configs:
  config1  ./ndt -d 4            config2  hypercube -d 8
  config3  random -d 6 (n=40, the reference kd builder explodes beyond ~200)
  config4  balls -d 5            config5  mixed10d (our C twin of the 10-D YAML) and the YAML itself
"""
import math
from collections import namedtuple

# cam: camera.h:16-20 (0 NORMAL, 1 VR, 2 PANO); stereo: ndt.c:46-48 (0 MONO, 1 SIDE_SIDE_3D,
# 2 OVER_UNDER_3D, 3 ANAGLYPH_3D, 4 HIDEF_3D)
Case = namedtuple("Case", "key scene dims cfg frame w h cam stereo", defaults=(0, 0))
H_FOV, V_FOV = 2.0 * math.pi, math.pi          # the defaults of -V / -P (ndt.c:1425-1426)

CASES = [
    Case("config1_default4d", None, 4, None, 0, 160, 90),
    Case("default3d", None, 3, None, 7, 96, 54),
    Case("default5d_odd", None, 5, None, 37, 97, 53),
    Case("config2_hypercube8d", "hypercube", 8, None, 0, 96, 54),
    Case("hypercube5d_hcube", "hypercube", 5, "hcube", 3, 64, 36),
    Case("hypercube6d_walls", "hypercube", 6, "walls", 0, 64, 36),
    Case("hypercube_points6d", "hypercube-points", 6, None, 0, 96, 54),
    Case("config3_random6d", "random", 6, "40", 0, 96, 54),
    Case("config4_balls5d", "balls", 5, None, 2, 96, 54),
    Case("config5_mixed10d", "mixed10d", 10, None, 0, 96, 54),
    Case("mixed7d", "mixed10d", 7, None, 5, 96, 54),
    Case("mixed12d", "mixed10d", 12, None, 11, 64, 36),
    # SURVEY 8(f) rank 3, BASELINE config 5 as written: scenes loaded by the reference's scenes/yaml.c +
    # scene_read_yaml (scene.c:2090) over include/yaml_lite; cfg = the YAML file (-u), repo-relative.
    # config5_mixed10d.yaml is what `ndt -s mixed10d.so -d 10 -y` writes (scene_write_yaml, scene.c:1000);
    # handwritten4d.yaml is a 3-document (3-frame) file in README.md:292-431 style, frame 1 rendered.
    Case("config5_yaml10d", "yaml", 10, "tests/scenes/config5_mixed10d.yaml", 0, 96, 54),
    Case("yaml_handwritten4d", "yaml", 4, "tests/scenes/handwritten4d.yaml", 1, 96, 54),
]
# the other cameras and the stereo modes of render_pixel, same scenes
VIEW_CASES = [
    Case("view_sidebyside4d", None, 4, None, 0, 96, 54, 0, 1),
    Case("view_overunder7d", "mixed10d", 7, None, 5, 96, 54, 0, 2),
    Case("view_anaglyph5d", "hypercube", 5, "hcube", 3, 64, 36, 0, 3),
    Case("view_hidef4d", None, 4, None, 0, 24, 2205, 0, 4),
    Case("view_vr5d", "hypercube", 5, "hcube", 3, 96, 54, 1, 0),
    Case("view_pano7d", "mixed10d", 7, None, 5, 96, 54, 2, 0),
    Case("view_vr_sidebyside4d", None, 4, None, 0, 96, 54, 1, 1),
    Case("view_pano_anaglyph4d", None, 4, None, 0, 48, 28, 2, 3),
]
BASE_CASES = list(CASES)
CASES = CASES + VIEW_CASES
BY_KEY = {c.key: c for c in CASES}

# recursive (Whitted) anti-aliasing, -w / -a: (key, scene, dims, cfg, frame, w, h) and the
# (aa_diff, aa_depth) pairs rendered for each: the default -a, the two -q presets that refine, and the two
# that do not (ndt.c:1589-1624, 1411-1412)
AaCase = namedtuple("AaCase", "key scene dims cfg frame w h")
AA_CASES = [
    AaCase("aa_default4d", None, 4, None, 0, 96, 54),
    AaCase("aa_hcube5d", "hypercube", 5, "hcube", 3, 64, 36),
    AaCase("aa_mixed7d", "mixed10d", 7, None, 5, 40, 24),
]
AA_PARAMS = [(20, 4), (1, 2), (5, 0), (255, 0), (20, -1)]
