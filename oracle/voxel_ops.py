"""NumPy restatement of the reference's point<->voxel operators.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Semantics follow the reference's *GPU* (CuPy ElementwiseKernel) code paths, in
float32 and without FMA contraction, because those are the paths the
reference's pose pipeline actually runs; a ``numpy_semantics=True`` switch
reproduces the reference's ``forward_cpu`` where the two differ
(``np.round`` half-to-even vs C ``round`` half-away; ``floor`` vs ``(int)``).

Reference files (relative to /root/reference/morefusion/functions/geometry):
  average_voxelization_3d.py, max_voxelization_3d.py, voxelization_3d.py,
  interpolate_voxel_grid.py, truncated_distance_function.py,
  occupancy_grid_3d.py
"""

import numpy as np

F32 = np.float32
I32 = np.int32


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def c_roundf(x):
    """C ``roundf``: round half away from zero, exact in float32."""
    x = np.asarray(x, dtype=F32)
    t = np.trunc(x)
    frac = x - t  # exact
    return (t + np.where(np.abs(frac) >= F32(0.5), np.sign(x), F32(0))).astype(F32)


def f2i(x):
    """CUDA ``static_cast<int>(float)``: truncate, saturate, NaN -> 0."""
    x = np.asarray(x, dtype=F32)
    y = np.where(np.isnan(x), F32(0), x)
    y = np.clip(np.trunc(y).astype(np.float64), -2147483648.0, 2147483647.0)
    return y.astype(np.int64).astype(I32)


def voxel_index(points, origin, pitch, numpy_semantics=False):
    """idx = (int)round((p - origin) / pitch), per axis.

    average_voxelization_3d.py:84-86 (GPU) / :29 (CPU).
    """
    points = np.asarray(points, dtype=F32)
    origin = np.asarray(origin, dtype=F32).reshape(3)
    pitch = F32(pitch)
    f = (points - origin[None, :]) / pitch
    if numpy_semantics:
        return np.round(f).astype(np.int64).astype(I32)
    return f2i(c_roundf(f))


def _inbounds(idx, dims):
    X, Y, Z = dims
    return (
        (idx[:, 0] >= 0) & (idx[:, 0] < X)
        & (idx[:, 1] >= 0) & (idx[:, 1] < Y)
        & (idx[:, 2] >= 0) & (idx[:, 2] < Z)
    )


def _check_dimensions(dimensions):
    # voxelization_3d.py:11-16
    if not (
        isinstance(dimensions, tuple)
        and len(dimensions) == 3
        and all(isinstance(d, int) for d in dimensions)
    ):
        raise ValueError("dimensions must be a tuple of 4 integers")


# --------------------------------------------------------------------------
# a1  average_voxelization_3d      (average_voxelization_3d.py:8-115)
# --------------------------------------------------------------------------
def average_voxelization_3d_fwd(
    values, points, batch_indices, *, batch_size, origin, pitch, dimensions,
    numpy_semantics=False,
):
    """Returns (matrix [B,C,X,Y,Z] f32, counts [B,X,Y,Z] i32).

    fp32 sums are taken in ascending point order (the reference's CPU loop,
    :24-34; the GPU path's atomics are order-nondeterministic), then divided
    by the count (:113-115).
    """
    _check_dimensions(dimensions)
    values = np.asarray(values, dtype=F32)
    points = np.asarray(points, dtype=F32)
    batch_indices = np.asarray(batch_indices, dtype=I32)
    if np.isnan(points).sum():
        raise ValueError("points include nan")  # :13-14, :47-48
    B, C = int(batch_size), values.shape[1]
    X, Y, Z = dimensions
    V = X * Y * Z
    idx = voxel_index(points, origin, pitch, numpy_semantics)
    ok = _inbounds(idx, dimensions)
    flat = (idx[:, 0].astype(np.int64) * Y + idx[:, 1]) * Z + idx[:, 2]
    key = batch_indices.astype(np.int64) * V + flat
    key = key[ok]
    matrix = np.zeros((B, V, C), dtype=F32)  # voxel-major scratch
    counts = np.zeros((B * V,), dtype=I32)
    np.add.at(counts, key, 1)
    # np.add.at is unbuffered and processes indices in order -> ascending
    # point order fp32 accumulation, identical to the reference's CPU loop
    np.add.at(matrix.reshape(B * V, C), key, values[ok])
    nz = counts > 0
    m2 = matrix.reshape(B * V, C)
    m2[nz] = (m2[nz] / counts[nz][:, None].astype(F32)).astype(F32)
    matrix = np.ascontiguousarray(
        matrix.reshape(B, X, Y, Z, C).transpose(0, 4, 1, 2, 3)
    )
    return matrix, counts.reshape(B, X, Y, Z)


def average_voxelization_3d_bwd(
    gmatrix, counts, points, batch_indices, *, origin, pitch, dimensions,
    numpy_semantics=False,
):
    """gvalues[n,c] = gmatrix[b,c,idx]/counts[b,idx] (:170-206)."""
    gmatrix = np.asarray(gmatrix, dtype=F32)
    B, C = gmatrix.shape[:2]
    X, Y, Z = dimensions
    idx = voxel_index(points, origin, pitch, numpy_semantics)
    ok = _inbounds(idx, dimensions)
    P = points.shape[0]
    gvalues = np.zeros((P, C), dtype=F32)
    b = np.asarray(batch_indices)[ok]
    i = idx[ok]
    g = gmatrix[b, :, i[:, 0], i[:, 1], i[:, 2]]
    c = counts[b, i[:, 0], i[:, 1], i[:, 2]].astype(F32)
    gvalues[ok] = g / c[:, None]
    return gvalues


# --------------------------------------------------------------------------
# a6  max_voxelization_3d        (max_voxelization_3d.py:8-183)
# --------------------------------------------------------------------------
def max_voxelization_3d_fwd(
    values, points, batch_indices, intensities, *, batch_size, origin, pitch,
    dimensions, numpy_semantics=False,
):
    """Returns (matrix [B,C,X,Y,Z], indices [B,X,Y,Z] i32).

    Winner per voxel = the point with the largest intensity, the lowest point
    id among exact ties (the CPU loop :25-40 replaces only on strictly
    greater; the GPU path :102-118 is racy on ties).
    """
    _check_dimensions(dimensions)
    values = np.asarray(values, dtype=F32)
    points = np.asarray(points, dtype=F32)
    if np.isnan(points).sum():
        raise ValueError("points include nan")
    B, C = int(batch_size), values.shape[1]
    X, Y, Z = dimensions
    V = X * Y * Z
    idx = voxel_index(points, origin, pitch, numpy_semantics)
    ok = _inbounds(idx, dimensions)
    flat = (idx[:, 0].astype(np.int64) * Y + idx[:, 1]) * Z + idx[:, 2]
    key = np.asarray(batch_indices).astype(np.int64) * V + flat
    indices = np.full((B * V,), -1, dtype=I32)
    best = np.zeros((B * V,), dtype=F32)
    for n in np.nonzero(ok)[0]:
        k = key[n]
        if indices[k] < 0 or intensities[n] > best[k]:
            indices[k] = n
            best[k] = intensities[n]
    matrix = np.zeros((B * V, C), dtype=F32)
    hit = indices >= 0
    matrix[hit] = values[indices[hit]]
    matrix = np.ascontiguousarray(
        matrix.reshape(B, X, Y, Z, C).transpose(0, 4, 1, 2, 3)
    )
    return matrix, indices.reshape(B, X, Y, Z)


def max_voxelization_3d_bwd(gmatrix, indices, n_points):
    """gvalues[n] = sum over voxels won by n of gmatrix[b,:,v] (:158-179)."""
    gmatrix = np.asarray(gmatrix, dtype=F32)
    B, C = gmatrix.shape[:2]
    g = gmatrix.reshape(B, C, -1).transpose(0, 2, 1).reshape(-1, C)
    ind = indices.reshape(-1)
    gvalues = np.zeros((n_points, C), dtype=F32)
    hit = ind >= 0
    np.add.at(gvalues, ind[hit], g[hit])
    return gvalues


# --------------------------------------------------------------------------
# a5  interpolate_voxel_grid     (interpolate_voxel_grid.py:6-59, 170-266)
# --------------------------------------------------------------------------
# corner order w000,w100,w010,w001,w110,w011,w101,w111 (:25-32)
_CORNERS = np.array(
    [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1],
     [1, 1, 0], [0, 1, 1], [1, 0, 1], [1, 1, 1]], dtype=I32)


def trilinear_params(points, numpy_semantics=False):
    """weights [P,8] f32, corner indices [P,8,3] i32 (:6-59 GPU, :62-113 CPU)."""
    p = np.asarray(points, dtype=F32)
    if numpy_semantics:
        lo = np.floor(p).astype(I32)
    else:
        lo = f2i(p)  # static_cast<int>: truncation toward zero
    if numpy_semantics:
        # f32 scalar - i32 scalar promotes to float64 in the reference's CPU
        # helper (:67-72); weights are rounded to f32 only when stored (:74-82)
        lw = p.astype(np.float64) - lo.astype(np.float64)
        hw = 1.0 - lw
    else:
        lw = (p - lo.astype(F32)).astype(F32)       # lx,ly,lz
        hw = (F32(1.0) - lw).astype(F32)            # hx,hy,hz
    w = np.empty((p.shape[0], 8), dtype=F32)
    for j, (cx, cy, cz) in enumerate(_CORNERS):
        wx = lw[:, 0] if cx else hw[:, 0]
        wy = lw[:, 1] if cy else hw[:, 1]
        wz = lw[:, 2] if cz else hw[:, 2]
        w[:, j] = (wx * wy).astype(lw.dtype) * wz   # (a*b)*c, as written
    ixyz = lo[:, None, :] + _CORNERS[None, :, :]
    return w, ixyz


def interpolate_voxel_grid_fwd(voxelized, points, batch_indices,
                               numpy_semantics=False):
    """values [P,C]; corners accumulated in order j=0..7 (:188-208).

    Uses the correct (Y*Z, Z, 1) strides; the reference forward kernel's
    (X*Y, Y, 1) (:201-205) is identical for cubic grids, to which parity
    tests are restricted.
    """
    vox = np.asarray(voxelized, dtype=F32)
    B, C, X, Y, Z = vox.shape
    w, ixyz = trilinear_params(points, numpy_semantics)
    P = w.shape[0]
    out = np.zeros((P, C), dtype=F32)
    b = np.asarray(batch_indices)
    for j in range(8):
        ix, iy, iz = ixyz[:, j, 0], ixyz[:, j, 1], ixyz[:, j, 2]
        ok = (ix >= 0) & (ix < X) & (iy >= 0) & (iy < Y) & (iz >= 0) & (iz < Z)
        v = vox[b[ok], :, ix[ok], iy[ok], iz[ok]]
        out[ok] = out[ok] + (w[ok, j][:, None] * v).astype(F32)
    return out


def interpolate_voxel_grid_bwd(gvalues, points, batch_indices, shape,
                               numpy_semantics=False):
    """gvoxelized [B,C,X,Y,Z] (:236-262); accumulation in (point, corner) order."""
    B, C, X, Y, Z = shape
    g = np.asarray(gvalues, dtype=F32)
    w, ixyz = trilinear_params(points, numpy_semantics)
    out = np.zeros((B, X, Y, Z, C), dtype=F32)
    b = np.asarray(batch_indices).astype(np.int64)
    P = g.shape[0]
    pid = np.repeat(np.arange(P), 8)
    ix = ixyz[:, :, 0].reshape(-1)
    iy = ixyz[:, :, 1].reshape(-1)
    iz = ixyz[:, :, 2].reshape(-1)
    ok = (ix >= 0) & (ix < X) & (iy >= 0) & (iy < Y) & (iz >= 0) & (iz < Z)
    flat = ((b[pid] * X + ix) * Y + iy) * Z + iz
    contrib = (w.reshape(-1)[:, None] * g[pid]).astype(F32)
    np.add.at(out.reshape(-1, C), flat[ok], contrib[ok])
    return np.ascontiguousarray(out.transpose(0, 4, 1, 2, 3))


# --------------------------------------------------------------------------
# a3  truncated_distance_function  (truncated_distance_function.py:21-166)
# --------------------------------------------------------------------------
def tdf_ksize(pitch, truncation):
    """ksize = ceil(trunc/pitch), made odd (:36-38); fp32 division."""
    k = int(np.ceil(F32(truncation) / F32(pitch)))
    if k % 2 == 0:
        k += 1
    return k


def tdf_kernel_offsets(ksize):
    """The reference's offset table (:39-41): meshgrid(indexing='xy')."""
    a = np.arange(ksize)
    k = np.stack(np.meshgrid(a, a, a), -1).reshape(-1, 3).astype(F32)
    return k - F32(ksize // 2)


def truncated_distance_function_fwd(points, *, pitch, origin, dims, truncation):
    """Returns (tdf [X,Y,Z] f32, indices [X,Y,Z] i32 = winning *point* id or -1).

    :51-79.  tdf[v] = min(trunc, min_p pitch*||f_p - v||) over points whose
    rounded voxel is within the ksize cube of v and whose distance < trunc.
    Winner = lowest flat thread id p*K+k among exact minima (the reference's
    atomicMin/atomicExch pair is racy on ties; serial execution of its kernel
    yields exactly this choice), reported as ``indices // K`` (:177).
    """
    p = np.asarray(points, dtype=F32)
    pitch = F32(pitch)
    origin = np.asarray(origin, dtype=F32).reshape(3)
    trunc = F32(truncation)
    X, Y, Z = dims
    V = X * Y * Z
    ks = tdf_ksize(pitch, trunc)
    offs = tdf_kernel_offsets(ks)                     # [K,3] float
    K = offs.shape[0]
    P = p.shape[0]
    f = ((p - origin[None]) / pitch).astype(F32)      # [P,3]
    r = c_roundf(f)                                   # [P,3] float
    tdf = np.full((V,), trunc, dtype=F32)
    ind = np.full((V,), -1, dtype=np.int64)
    if P == 0:
        return tdf.reshape(dims), ind.astype(I32).reshape(dims)
    # int ix = round(ix_f) + kernel[3k]  (float add, then (int))
    vox = f2i((r[:, None, :] + offs[None, :, :]).astype(F32))       # [P,K,3]
    d = (f[:, None, :] - vox.astype(F32)).astype(F32)               # [P,K,3]
    s = ((d[..., 0] * d[..., 0]).astype(F32) + (d[..., 1] * d[..., 1]).astype(F32)).astype(F32)
    s = (s + (d[..., 2] * d[..., 2]).astype(F32)).astype(F32)
    dist = (pitch * np.sqrt(s).astype(F32)).astype(F32)             # [P,K]
    ok = (
        (vox[..., 0] >= 0) & (vox[..., 0] < X)
        & (vox[..., 1] >= 0) & (vox[..., 1] < Y)
        & (vox[..., 2] >= 0) & (vox[..., 2] < Z)
        & (dist < trunc)
    )
    flat = (vox[..., 0].astype(np.int64) * Y + vox[..., 1]) * Z + vox[..., 2]
    tid = np.arange(P * K, dtype=np.int64).reshape(P, K)
    fl, di, ti = flat[ok], dist[ok], tid[ok]
    # lexicographic min of (dist, tid) per voxel
    order = np.lexsort((ti, di, fl))
    fl, di, ti = fl[order], di[order], ti[order]
    first = np.ones(fl.shape[0], dtype=bool)
    first[1:] = fl[1:] != fl[:-1]
    tdf[fl[first]] = di[first]
    ind[fl[first]] = ti[first] // K
    return tdf.reshape(dims), ind.astype(I32).reshape(dims)


def truncated_distance_function_bwd(gtdf, points, indices, *, pitch, origin, dims):
    """gpoints [P,3] (:119-145): unit vector voxel->point times gtdf, summed
    over the voxels a point wins (ascending voxel order)."""
    p = np.asarray(points, dtype=F32)
    pitch = F32(pitch)
    origin = np.asarray(origin, dtype=F32).reshape(3)
    X, Y, Z = dims
    g = np.asarray(gtdf, dtype=F32).reshape(-1)
    ind = np.asarray(indices).reshape(-1)
    hit = np.nonzero(ind >= 0)[0]
    gp = np.zeros_like(p)
    if hit.size == 0:
        return gp
    pid = ind[hit]
    vx = (hit // (Y * Z)).astype(F32)
    vy = ((hit // Z) % Y).astype(F32)
    vz = (hit % Z).astype(F32)
    f = ((p[pid] - origin[None]) / pitch).astype(F32)
    d = np.stack([f[:, 0] - vx, f[:, 1] - vy, f[:, 2] - vz], 1).astype(F32)
    s = ((d[:, 0] * d[:, 0]).astype(F32) + (d[:, 1] * d[:, 1]).astype(F32)).astype(F32)
    s = (s + (d[:, 2] * d[:, 2]).astype(F32)).astype(F32)
    n = np.sqrt(s).astype(F32)
    keep = n > 0
    with np.errstate(invalid="ignore", divide="ignore"):
        u = ((d / n[:, None]).astype(F32) * g[hit][:, None]).astype(F32)
    np.add.at(gp, pid[keep], u[keep])
    return gp


# --------------------------------------------------------------------------
# a4  pseudo_occupancy_voxelization  (truncated_distance_function.py:181-213)
# --------------------------------------------------------------------------
def pseudo_occupancy_voxelization_fwd(
    points, sdf, *, pitch, origin, dims, threshold=1, sdf_offset=0,
):
    """Returns dict(grid, surface, inside, tdf, indices, w_surface, w_inside)."""
    pitch = F32(pitch)
    trunc = F32(F32(threshold) * pitch)
    tdf, ind = truncated_distance_function_fwd(
        points, pitch=pitch, origin=origin, dims=dims, truncation=trunc)
    grid = (F32(1) - (tdf / trunc).astype(F32)).astype(F32)
    sdf = np.asarray(sdf, dtype=F32)
    w_in = np.full(tdf.shape, -1, dtype=F32)
    m = ind != -1
    w_in[m] = sdf[ind[m]]
    w_in = (w_in + F32(sdf_offset)).astype(F32)
    mask = w_in < 0
    w_in[mask] = 0
    with np.errstate(invalid="ignore", divide="ignore"):
        w_in = (w_in / w_in.max()).astype(F32)        # 0/0 -> NaN by design
    w_surf = w_in.copy()
    w_surf[~mask] = F32(1) - w_surf[~mask]
    return dict(
        grid=grid, surface=(grid * w_surf).astype(F32),
        inside=(grid * w_in).astype(F32), tdf=tdf, indices=ind,
        w_surface=w_surf, w_inside=w_in, truncation=trunc,
    )


# --------------------------------------------------------------------------
# a2  occupancy_grid_3d            (occupancy_grid_3d.py:31-85)
# --------------------------------------------------------------------------
def occupancy_grid_3d_fwd(points, *, pitch, origin, dims, threshold=1):
    """m = min(relu(threshold - min_p ||ijk - (p-o)/pitch||), 1).

    Returns (m [X,Y,Z] f32, argmin bookkeeping for bwd)."""
    p = np.asarray(points, dtype=F32)
    pitch = F32(pitch)
    origin = np.asarray(origin, dtype=F32).reshape(3)
    X, Y, Z = [int(d) for d in dims]
    q = ((p - origin[None]) / pitch).astype(F32)
    I, J, K = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij")
    d0 = I[..., None].astype(F32) - q[None, None, None, :, 0]
    d1 = J[..., None].astype(F32) - q[None, None, None, :, 1]
    d2 = K[..., None].astype(F32) - q[None, None, None, :, 2]
    s = ((d0 * d0).astype(F32) + (d1 * d1).astype(F32)).astype(F32)
    s = (s + (d2 * d2).astype(F32)).astype(F32)
    d = np.sqrt(s).astype(F32)                         # [X,Y,Z,P]
    dmin = d.min(axis=3)
    m = np.maximum(F32(threshold) - dmin, F32(0)).astype(F32)
    m = np.minimum(m, F32(1))
    return m, dict(d=d, dmin=dmin, d0=d0, d1=d1, d2=d2)


def occupancy_grid_3d_bwd(gm, aux, *, pitch, threshold=1):
    """Chain rule of the chainer generic ops the reference composes (:77-85):
    F.minimum(m,1) passes grad where m<=1; F.relu where threshold-d>0; F.min
    routes to ALL tied minima; sqrt'(s)=1/(2 sqrt s); -1/pitch to points
    (:56-74)."""
    pitch = F32(pitch)
    d, dmin = aux["d"], aux["dmin"]
    g = np.asarray(gm, dtype=F32)
    r = F32(threshold) - dmin
    g = np.where(r > 0, g, F32(0))
    g = np.where(np.maximum(r, 0) <= 1, g, F32(0))
    gd = -g                                            # d(threshold - d)
    sel = d == dmin[..., None]
    gdp = np.where(sel, gd[..., None], F32(0))         # [X,Y,Z,P]
    with np.errstate(invalid="ignore", divide="ignore"):
        inv = gdp / d
    gx = -(inv * aux["d0"] / pitch).sum(axis=(0, 1, 2))
    gy = -(inv * aux["d1"] / pitch).sum(axis=(0, 1, 2))
    gz = -(inv * aux["d2"] / pitch).sum(axis=(0, 1, 2))
    return np.stack([gx, gy, gz], 1).astype(F32)
