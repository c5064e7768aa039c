"""Scene catalogue shared by the tests, the golden generator and bench.py.

Each entry names a reference scene plugin (oracle/_ref/scenes/<name>.so; None =
the built-in scene of scene.c:429-571), the dimension, the scene's config
string (-u), the frame, and a SMALL frame size at which the CPU oracle
finishes in about a second.  `cfg` rows map to BASELINE.json configs:
  config1  ./ndt -d 4            config2  hypercube -d 8
  config3  random -d 6 (n=40, the reference kd builder explodes beyond ~200)
  config4  balls -d 5            config5  mixed10d (our C twin of the 10-D YAML)
"""
from collections import namedtuple

Case = namedtuple("Case", "key scene dims cfg frame w h")

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
]
BY_KEY = {c.key: c for c in CASES}
