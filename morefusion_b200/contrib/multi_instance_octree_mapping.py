"""Occupancy-grid producer on the GPU.

API of morefusion/contrib/multi_instance_octree_mapping.py:7-133 -- ``MultiInstanceOctreeMapping``
with ``instance_ids``, ``initialize(instance_id, *, pitch)``, ``integrate(instance_id, mask, pcd,
origin=(0, 0, 0))``, ``update(instance_id, occupied)``, ``get_target_grids(target_id, *,
dimensions, pitch, origin)`` and ``get_target_pcds(target_id, aabb_min=None, aabb_max=None)`` --
which is what builds ``grid_target / grid_nontarget / grid_empty`` for the pose network and the ICC
(datasets/rgbd_pose_estimation/base.py:30-50,157-162).

The reference keeps one OctoMap ``OcTree`` per instance on the host and queries it voxel by voxel
from Python (32 768 ``search`` calls per instance and target).  Here every instance lives in one
device-resident hash table of log-odds cells (csrc/mapping.cu): a depth scan is two launches, the
grids of any number of targets are one launch (``get_target_grids_batch`` returns CUDA tensors for
the per-frame path; ``get_target_grids`` returns numpy arrays like the reference).  OctoMap's
sensor model and key arithmetic are kept (hit 0.7, miss 0.4, clamping [0.1192, 0.971], 16-bit keys);
the library itself is absent from the reference tree, see oracle/octomap.py for the statement of
what is reproduced.  No CPU path: tensors live on a CUDA device.
"""

import math

import numpy as np
import torch

from .. import _lib

_TREE_MAX_VAL = 32768


def _logodds(p):
    return np.float32(math.log(p / (1.0 - p)))


class MultiInstanceOctreeMapping:
    PROB_HIT, PROB_MISS, CLAMP_MIN, CLAMP_MAX = 0.7, 0.4, 0.1192, 0.971     # OctoMap defaults

    def __init__(self, device=None, capacity=1 << 22):
        """capacity: initial number of hash slots (power of two, 20 bytes each; the table grows x4
        when a quarter full)."""
        if capacity < 64 or capacity & (capacity - 1):
            raise ValueError("capacity must be a power of two >= 64")
        self.device = torch.device("cuda" if device is None else device)
        self._ids = {}                           # instance_id -> dense index (insertion order)
        self._pitch = []                         # per dense index
        self._scan = 0
        self._cap = int(capacity)
        self._hit, self._miss = _logodds(self.PROB_HIT), _logodds(self.PROB_MISS)
        self._lo_min, self._lo_max = _logodds(self.CLAMP_MIN), _logodds(self.CLAMP_MAX)
        self._alloc_table(self._cap)
        self._counters = torch.zeros(8, dtype=torch.int32, device=self.device)
        self._host_counters = None
        self._pending = None                     # event of the last asynchronous counter read-back
        self._res_factor = None                  # cached device array of 1 / pitch
        self._lut = None                         # cached (lut, lut_lo, resolutions) device arrays
        self._query_cache = None                 # cached (key, targets, pitches, origins) device arrays

    # ------------------------------------------------------------------ table management
    def _alloc_table(self, cap):
        dev = self.device
        # 16-byte cells {u64 key | f32 log-odds, u32 stamp}: column 0 = key (-1 = free slot),
        # column 1 = the value / stamp pair
        self._cells = torch.zeros((cap, 2), dtype=torch.int64, device=dev)
        self._cells[:, 0] = -1
        self._cnt = torch.zeros(cap, dtype=torch.int32, device=dev)
        self._cap = cap

    @property
    def _keys(self):
        return self._cells[:, 0]

    @property
    def _lo(self):
        return self._cells.view(torch.float32)[:, 2]

    def _table_args(self):
        p = _lib.ptr
        return (p(self._cells), p(self._cnt), self._cap, p(self._counters))

    def _read_back(self):
        """Asynchronous copy of the counters into pinned host memory; looked at by the next call."""
        if self._host_counters is None:
            self._host_counters = torch.zeros(8, dtype=torch.int32).pin_memory()
        self._host_counters.copy_(self._counters, non_blocking=True)
        self._pending = torch.cuda.Event()
        self._pending.record(torch.cuda.current_stream(self.device))

    def _check(self, wait=True):
        """Act on the counters of the previous operation: errors for dropped work, growth.
        `wait=False` (the per-frame calls integrate_labels / get_target_grids_batch) never blocks
        the host: if the read-back of the previous operation has not landed yet it is looked at by
        a later call (the table grows at a quarter full, so a decision that lags a scan is still
        early; overflow is sticky and cannot be missed).  The reference-style calls wait, which
        keeps growth decisions exact for small tables."""
        if self._pending is None:
            return
        if not wait and not self._pending.query():
            return
        self._pending.synchronize()
        self._pending = None
        c = self._host_counters.tolist()
        if c[1]:
            raise RuntimeError(
                f"occupancy map: hash table of {self._cap} slots overflowed during the previous "
                "operation and updates were dropped; construct the mapping with a larger capacity")
        if c[0] * 4 > self._cap:
            self._grow(self._cap * 4)

    def _grow(self, new_cap):
        L = _lib.lib()
        old = (self._cells, self._cap)
        self._alloc_table(new_cap)
        with self._dev_ctx():
            _lib.check(L.mf_map_rehash(_lib.ptr(old[0]), old[1], *self._table_args(), _lib.stream()),
                       "map_rehash")

    def _dev_ctx(self):
        return torch.cuda.device(self.device)

    # ------------------------------------------------------------------ reference API
    @property
    def instance_ids(self):
        return list(self._ids.keys())

    def initialize(self, instance_id, *, pitch):
        if instance_id in self._ids:
            raise ValueError("instance {instance_id} already exists")
        if not pitch > 0:
            raise ValueError("pitch must be positive")
        if len(self._ids) >= 0xFFFF:
            raise ValueError("too many instances")
        self._ids[instance_id] = len(self._ids)
        self._pitch.append(float(pitch))
        self._res_factor = None
        self._lut = None

    def integrate(self, instance_id, mask, pcd, origin=(0, 0, 0)):
        """octree.insertPointCloud(pcd[mask & nonnan], origin) (:21-27).  mask [H,W] bool, pcd
        [H,W,3]; numpy arrays or tensors (CUDA tensors are used in place)."""
        idx = self._ids[instance_id]                       # KeyError for an unknown instance, as the dict
        dev = self.device
        pts = torch.as_tensor(pcd).to(device=dev, dtype=torch.float32).contiguous()
        msk = torch.as_tensor(mask).to(device=dev)
        if pts.dim() < 2 or pts.shape[-1] != 3 or tuple(msk.shape) != tuple(pts.shape[:-1]):
            raise ValueError("pcd must be [..., 3] and mask its leading shape")
        if msk.dtype != torch.bool:
            raise ValueError("mask must be bool")
        _lib.require_cuda(pts, msk)
        msk = msk.contiguous().view(torch.uint8)
        n = pts.numel() // 3
        org = np.asarray(origin, dtype=np.float64).astype(np.float32)      # point3d: floats
        self._check()
        if n == 0:
            return
        self._scan += 1
        with self._dev_ctx():
            _lib.check(_lib.lib().mf_map_integrate(
                _lib.ptr(pts), _lib.ptr(msk), n, float(org[0]), float(org[1]), float(org[2]),
                self._pitch[idx], idx, self._scan, float(self._hit), float(self._miss),
                float(self._lo_min), float(self._lo_max), *self._table_args(), _lib.stream()),
                "map_integrate")
            self._read_back()

    def integrate_labels(self, label, pcd, origin=(0, 0, 0)):
        """Every initialised instance of a labelled frame in one scan: equal to
        ``for i in instance_ids: integrate(i, label == i, pcd, origin)`` (what build_octomap,
        datasets/rgbd_pose_estimation/base.py:30-50, does instance by instance), as two launches.
        label [H,W] integer image; pixels whose label is not an initialised instance are skipped."""
        dev = self.device
        pts = torch.as_tensor(pcd).to(device=dev, dtype=torch.float32).contiguous()
        lab = torch.as_tensor(label).to(device=dev)
        if pts.dim() < 2 or pts.shape[-1] != 3 or tuple(lab.shape) != tuple(pts.shape[:-1]):
            raise ValueError("pcd must be [..., 3] and label its leading shape")
        if lab.dtype.is_floating_point or lab.dtype == torch.bool:
            raise ValueError("label must be an integer image")
        _lib.require_cuda(pts, lab)
        lab = lab.to(torch.int32).contiguous()
        n = pts.numel() // 3
        org = np.asarray(origin, dtype=np.float64).astype(np.float32)
        self._check(wait=False)
        if n == 0 or not self._ids:
            return
        if self._lut is None:
            ids = [int(i) for i in self._ids]            # instance ids must be integers here
            lo, hi = min(ids), max(ids)
            if hi - lo >= 1 << 20:
                raise ValueError("instance ids span too wide a range for a label look-up table")
            lut = np.full(hi - lo + 1, -1, np.int32)
            for i, idx in self._ids.items():
                lut[int(i) - lo] = idx
            self._lut = (torch.as_tensor(lut).to(dev), lo,
                         torch.as_tensor(np.asarray(self._pitch, dtype=np.float64)).to(dev))
        lut, lo, res = self._lut
        self._scan += 1
        with self._dev_ctx():
            _lib.check(_lib.lib().mf_map_integrate_labelled(
                _lib.ptr(pts), _lib.ptr(lab), n, float(org[0]), float(org[1]), float(org[2]),
                _lib.ptr(lut), lo, lut.numel(), _lib.ptr(res), self._scan, float(self._hit),
                float(self._miss), float(self._lo_min), float(self._lo_max), *self._table_args(),
                _lib.stream()), "map_integrate_labelled")
            self._read_back()

    def update(self, instance_id, occupied):
        """octree.updateNodes(occupied, True) (:29-34): one hit update per row of occupied [M,3]."""
        idx = self._ids[instance_id]
        pts = torch.as_tensor(occupied).to(device=self.device, dtype=torch.float64).reshape(-1, 3).contiguous()
        _lib.require_cuda(pts)
        self._check()
        if pts.shape[0] == 0:
            return
        with self._dev_ctx():
            _lib.check(_lib.lib().mf_map_update_points(
                _lib.ptr(pts), pts.shape[0], self._pitch[idx], idx, float(self._hit),
                float(self._lo_min), float(self._lo_max), *self._table_args(), _lib.stream()),
                "map_update_points")
            self._read_back()

    def get_target_grids_batch(self, target_ids, *, dimensions, pitches, origins):
        """The three grids of several targets from one launch.  Returns CUDA tensors
        (grid_target, grid_nontarget, grid_empty), each [T, X, Y, Z] float32."""
        assert len(dimensions) == 3
        assert (np.asarray(dimensions) > 0).all()
        T = len(target_ids)
        pit = np.asarray(pitches, dtype=np.float64).reshape(T)
        org = np.asarray(origins, dtype=np.float64).reshape(T, 3)
        assert not np.isnan(org).any()
        assert (pit > 0).all()
        tix = np.asarray([self._ids[t] for t in target_ids], dtype=np.int32)
        dev = self.device
        self._check(wait=False)
        if self._res_factor is None:
            self._res_factor = torch.as_tensor(1.0 / np.asarray(self._pitch, dtype=np.float64)).to(dev)
        X, Y, Z = (int(d) for d in dimensions)
        out = torch.empty((3, T, X, Y, Z), dtype=torch.float32, device=dev)
        # the per-frame path asks for the same targets frame after frame: their small device
        # arrays are uploaded once per distinct (targets, pitches, origins)
        key = (tix.tobytes(), pit.tobytes(), org.tobytes())
        if self._query_cache is None or self._query_cache[0] != key:
            self._query_cache = (key, torch.as_tensor(tix).to(dev), torch.as_tensor(pit).to(dev),
                                 torch.as_tensor(org).to(dev))
        _, d_tix, d_pit, d_org = self._query_cache
        with self._dev_ctx():
            _lib.check(_lib.lib().mf_map_query_grids(
                _lib.ptr(d_tix), _lib.ptr(d_pit), _lib.ptr(d_org), T, X, Y, Z,
                _lib.ptr(self._res_factor), len(self._pitch), *self._table_args(),
                _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.ptr(out[2]), _lib.stream()), "map_query_grids")
        return out[0], out[1], out[2]

    def get_target_grids(self, target_id, *, dimensions, pitch, origin):
        """(grid_target, grid_nontarget, grid_empty) as float32 numpy arrays of shape `dimensions`
        (:35-94)."""
        assert not np.isnan(origin).any()
        assert pitch > 0
        self._check()
        gt, gn, ge = self.get_target_grids_batch(
            [target_id], dimensions=dimensions, pitches=[pitch], origins=[origin])
        return gt[0].cpu().numpy(), gn[0].cpu().numpy(), ge[0].cpu().numpy()

    def get_target_pcds(self, target_id, aabb_min=None, aabb_max=None):
        """Centres of the occupied / free cells of one instance, float64 [N,3] / [M,3] (:96-133).
        Rows are in key order (the reference returns them in OctoMap's tree-iteration order)."""
        idx = self._ids[target_id]
        self._check()
        keys = self._keys
        sel = (keys != -1) & (((keys >> 48) & 0xFFFF) == idx)
        k = keys[sel]
        lo = self._lo[sel]
        order = torch.argsort(k)
        k, lo = k[order], lo[order]
        kxyz = torch.stack([(k >> 32) & 0xFFFF, (k >> 16) & 0xFFFF, k & 0xFFFF], 1)
        coords = ((kxyz - _TREE_MAX_VAL).to(torch.float64) + 0.5) * self._pitch[idx]
        coords = coords.to(torch.float32).to(torch.float64).cpu().numpy()   # point3d floats -> float64
        occ = (lo >= 0).cpu().numpy()
        # axis-aligned box [aabb_min, aabb_max): lower bound inclusive, upper exclusive (:124-131)
        inside = np.ones(len(coords), dtype=bool)
        if aabb_min is not None:
            inside &= (coords >= np.asarray(aabb_min)).all(axis=1)
        if aabb_max is not None:
            inside &= (coords < np.asarray(aabb_max)).all(axis=1)
        return coords[occ & inside], coords[~occ & inside]

    # ------------------------------------------------------------------ introspection (tests, bench)
    def n_cells(self):
        return int(self._counters[0].item())

    def cells(self, instance_id):
        """{(kx, ky, kz): log-odds} of one instance (host dict; for tests)."""
        idx = self._ids[instance_id]
        cells = self._cells.cpu()
        keys = cells[:, 0].numpy()
        lo = cells.view(torch.float32)[:, 2].numpy()
        sel = (keys != -1) & (((keys >> 48) & 0xFFFF) == idx)
        out = {}
        for k, v in zip(keys[sel].tolist(), lo[sel].tolist()):
            out[((k >> 32) & 0xFFFF, (k >> 16) & 0xFFFF, k & 0xFFFF)] = np.float32(v)
        return out
