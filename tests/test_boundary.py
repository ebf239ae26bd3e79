"""CPU-only: the drop-in surface: the `morefusion` import alias resolves to the
B200 package, host-side metrics equal the reference's definitions."""

import subprocess
import sys

import numpy as np

from oracle import metrics as ometrics


def test_morefusion_alias_resolves():
    code = (
        "import morefusion, morefusion.functions as F, morefusion.contrib as C;"
        "from morefusion.contrib.singleview_3d.models import Model;"
        "import morefusion.metrics, morefusion.geometry;"
        "names = ['average_voxelization_3d','compose_transform','interpolate_voxel_grid',"
        "'max_voxelization_3d','occupancy_grid_3d','pseudo_occupancy_voxelization',"
        "'quaternion_matrix','transform_points','transformation_matrix','translation_matrix',"
        "'truncated_distance_function','average_distance'];"      # functions/__init__.py:3-15
        "assert all(hasattr(F, n) for n in names), [n for n in names if not hasattr(F, n)];"
        "assert hasattr(C, 'IterativeCollisionCheckLink') and hasattr(C, 'OccupancyRegistration');"
        "assert F is __import__('morefusion_b200').functions; print('ok')")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr[-2000:]


def test_ycb_video_add_auc_matches_reference_definition():
    from morefusion_b200.metrics import ycb_video_add_auc
    rs = np.random.RandomState(0)
    for _ in range(20):
        adds = rs.gamma(2.0, 0.03, size=rs.randint(1, 60))
        assert abs(ycb_video_add_auc(adds, max_value=0.1) - ometrics.ycb_video_add_auc(adds, 0.1)) < 1e-12
    assert ycb_video_add_auc(np.array([0.5, 0.7])) == 0            # all misses
    auc, x, y = ycb_video_add_auc(np.array([0.01, 0.05]), return_xy=True)
    assert x[0] == 0 and x[-1] == 0.1 and y[-1] == 1.0


def test_auc_for_errors_matches_reference_definition():
    import sklearn.metrics
    from morefusion_b200.metrics import auc_for_errors
    e = np.random.RandomState(1).rand(77) * 0.3
    x = np.linspace(0, 0.2, 1000)
    y = np.array([(e <= t).sum() / e.size for t in x])       # auc_for_errors.py:13-18
    assert abs(auc_for_errors(e, 0.2) - sklearn.metrics.auc(x, y) / 0.2) < 1e-12
