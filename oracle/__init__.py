"""CPU oracle for the MoreFusion volumetric-pose hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: it is a
NumPy / torch-CPU restatement of the reference's algorithms (each function
cites the reference file:line it follows) and may be imported only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- as the checker, never as the thing
measured or shipped.  ``morefusion_b200`` never imports it.

Pinning status (see DESIGN.md "Oracle"):
  * pinned by the reference's own code, executed here by
    ``oracle/ref_harness`` (reference ``forward_cpu``/``backward_cpu`` NumPy
    paths run verbatim under a chainer stub, and the reference's CuPy
    ``ElementwiseKernel`` source strings compiled as serial C++ into
    ``oracle/_ref/``): average_voxelization_3d, max_voxelization_3d,
    interpolate_voxel_grid, truncated_distance_function (fwd+bwd),
    pseudo_occupancy_voxelization, occupancy_grid_3d, quaternion_matrix,
    compose_transform, translation_matrix, transformation_matrix,
    transform_points, IterativeCollisionCheckLink.forward (loss value);
    golden vectors committed under ``tests/golden/``.
  * pinned by the reference's known-answer tests: occupancy_grid_3d
    (tests/functions_tests/geometry_tests/test_occupancy_grid_3d.py:24-38).
  * PARITY UNPINNED (third-party code absent from /root/reference):
    chainer.optimizers.Adam update rule, cuDNN ConvolutionND numerics,
    trimesh quaternion_from_matrix.  Restated from their published
    definitions; see ``oracle/icc.py`` and ``oracle/cnn.py`` headers.
    ``oracle/cnn_train.py`` (differentiable restatement of the same forward + the training
    loss, model.py:377-441) is pinned to ``oracle/cnn.py`` bit for bit on the forward pass and
    to finite differences on its gradients; the reference pins neither.
"""
