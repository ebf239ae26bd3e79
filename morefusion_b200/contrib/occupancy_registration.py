"""Single-object occupancy alignment on the CUDA operators.

API of morefusion/contrib/occupancy_registration.py:10-139: ``OccupancyRegistrationLink`` with
parameters ``quaternion`` [4] / ``translation`` [3] and ``forward(points_source, grid_target, *,
pitch, origin, threshold) -> loss`` (:21-59: transform the source points, matrix-free
``occupancy_grid_3d``, reward = overlap with the occupied target, penalty = overlap with the
unoccupied target); ``OccupancyRegistration(points_source, grid_target, *, pitch, origin,
threshold, transform_init, gpu=0, alpha=0.1)`` with ``register_iterative`` / ``register``
(:62-139) driving it with Chainer-form Adam (translation alpha x0.1)."""

import numpy as np
import torch

from .. import functions
from ..geometry import quaternion_from_matrix, translation_from_matrix
from ..optimizers import ChainerAdam


class OccupancyRegistrationLink(torch.nn.Module):
    def __init__(self, quaternion_init=None, translation_init=None):
        super().__init__()
        if quaternion_init is None:
            quaternion_init = np.array([1, 0, 0, 0], dtype=np.float32)
        if translation_init is None:
            translation_init = np.zeros((3,), dtype=np.float32)
        self.quaternion = torch.nn.Parameter(torch.as_tensor(np.asarray(quaternion_init, np.float32)))
        self.translation = torch.nn.Parameter(torch.as_tensor(np.asarray(translation_init, np.float32)))

    def forward(self, points_source, grid_target, *, pitch, origin, threshold):
        transform = functions.quaternion_matrix(self.quaternion[None])
        transform = functions.compose_transform(transform[:, :3, :3], self.translation[None])
        points_source = functions.transform_points(points_source, transform)[0]
        grid_source = functions.occupancy_grid_3d(
            points_source, pitch=pitch, origin=origin, dims=tuple(grid_target.shape[1:]),
            threshold=threshold)
        assert grid_target.dtype == torch.float32
        occupied_target = grid_target[0]
        reward = torch.sum(occupied_target * grid_source) / torch.sum(occupied_target)
        if grid_target.shape[0] == 3:
            unoccupied_target = torch.maximum(grid_target[1], grid_target[2])
        else:
            assert grid_target.shape[0] == 2
            unoccupied_target = grid_target[1]
        penalty = torch.sum(unoccupied_target * grid_source) / torch.sum(grid_source)
        return -reward + penalty


class OccupancyRegistration:
    def __init__(self, points_source, grid_target, *, pitch, origin, threshold, transform_init,
                 gpu=0, alpha=0.1):
        if gpu < 0:
            raise RuntimeError("OccupancyRegistration runs on CUDA only (no CPU fallback)")
        dev = torch.device("cuda", gpu)
        T = np.asarray(transform_init)
        link = OccupancyRegistrationLink(quaternion_from_matrix(T).astype(np.float32),
                                         translation_from_matrix(T).astype(np.float32)).to(dev)
        self._link = link
        self._grid_target_cpu = grid_target
        self._points_source = torch.as_tensor(np.asarray(points_source, np.float32), device=dev)
        self._grid_target = torch.as_tensor(np.asarray(grid_target, np.float32), device=dev)
        self._pitch = pitch
        self._origin = origin
        self._threshold = threshold
        self._optimizer = ChainerAdam([
            dict(params=[link.quaternion], alpha=alpha),
            dict(params=[link.translation], alpha=alpha * 0.1)])

    @property
    def _transform(self):
        with torch.no_grad():
            R = functions.quaternion_matrix(self._link.quaternion[None])
            T = functions.compose_transform(R[:, :3, :3], self._link.translation[None])
        return T[0].cpu().numpy()

    def register_iterative(self, iteration=None):
        iteration = 100 if iteration is None else iteration
        yield self._transform
        for _ in range(iteration):
            loss = self._link(points_source=self._points_source, grid_target=self._grid_target,
                              pitch=self._pitch, origin=self._origin, threshold=self._threshold)
            self._optimizer.zero_grad(set_to_none=True)
            loss.backward()
            self._optimizer.step()
            yield self._transform

    def register(self, iteration=None):
        for _ in self.register_iterative(iteration=iteration):
            pass
        return self._transform
